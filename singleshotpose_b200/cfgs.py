"""Programmatic generators for the two network description files the hot path is quoted on.

The reference ships them as text (``cfg/yolo-pose.cfg`` and
``multi_obj_pose_estimation/cfg/yolo-pose-multi.cfg``).  Those files do not exist on the
GPU box, so tests / bench / smoke synthesise an equivalent description from the compact
layer table below.  ``parse_cfg`` (cfg.py) accepts both these generated files and the
reference's own files unchanged.

Layer table follows reference cfg/yolo-pose.cfg:31-265 (Darknet-19 trunk, passthrough
route/reorg, 1x1 linear head, region block).
"""
from __future__ import annotations

import os
import tempfile

# (filters, size) for the batch-normalised leaky convs; 'M' = maxpool 2/2.
_TRUNK = [
    (32, 3), 'M', (64, 3), 'M', (128, 3), (64, 1), (128, 3), 'M',
    (256, 3), (128, 1), (256, 3), 'M',
    (512, 3), (256, 1), (512, 3), (256, 1), (512, 3), 'M',
    (1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3),
    (1024, 3), (1024, 3),
]

_MULTI_ANCHORS = "1.4820, 2.2412, 2.0501, 3.1265, 2.3946, 4.6891, 3.1018, 3.9910, 3.4879, 5.8851"


def _conv(filters, size, bn=True, act="leaky"):
    s = "[convolutional]\n"
    if bn:
        s += "batch_normalize=1\n"
    s += "filters=%d\nsize=%d\nstride=1\npad=1\nactivation=%s\n\n" % (filters, size, act)
    return s


def yolo_pose_cfg_text(multi: bool = False, width: int = 416, height: int = 416) -> str:
    """Text of yolo-pose.cfg (single object: 20 output channels, 1 anchor, 1 class) or
    yolo-pose-multi.cfg (160 channels, 5 anchors, 13 classes)."""
    net = ["[net]", "batch=%d" % (32 if multi else 8), "height=%d" % height, "width=%d" % width,
           "channels=3", "num_keypoints=9", "momentum=0.9", "decay=0.0005", "angle=0",
           "learning_rate=0.001", "burn_in=1000", "max_batches = 80200", "policy=steps",
           "max_epochs=500"]
    if multi:
        net += ["steps=-1,100,20000,30000", "scales=0.1,10,.1,.1", "conf_thresh = 0.05"]
    else:
        net += ["steps=-1,80,160", "scales=0.1,0.1,0.1", "conf_thresh= 0.1",
                "test_width=672", "test_height=672"]
    net += ["saturation = 1.5", "exposure = 1.5", "hue=.1", ""]
    out = "\n".join(net) + "\n"
    for item in _TRUNK:
        if item == 'M':
            out += "[maxpool]\nsize=2\nstride=2\n\n"
        else:
            out += _conv(item[0], item[1])
    out += "[route]\nlayers=-9\n\n"
    out += _conv(64, 1)
    out += "[reorg]\nstride=2\n\n"
    out += "[route]\nlayers=-1,-4\n\n"
    out += _conv(1024, 3)
    out += _conv(160 if multi else 20, 1, bn=False, act="linear")
    out += "[region]\nanchors = %s\nbias_match=1\nclasses=%d\ncoords=18\nnum=%d\n" % (
        _MULTI_ANCHORS if multi else "", 13 if multi else 1, 5 if multi else 1)
    out += ("softmax=1\njitter=.3\nrescore=1\n\nobject_scale=5\nnoobject_scale=0.1\n"
            "class_scale=1\ncoord_scale=1\n\nabsolute=1\nthresh = .6\nrandom=1\n")
    return out


def write_cfg(path: str | None = None, multi: bool = False, **kw) -> str:
    """Write the generated cfg to *path* (default: a temp file) and return the path."""
    if path is None:
        fd, path = tempfile.mkstemp(suffix="-multi.cfg" if multi else ".cfg", prefix="yolo-pose-")
        os.close(fd)
    with open(path, "w") as f:
        f.write(yolo_pose_cfg_text(multi=multi, **kw))
    return path


# Camera intrinsics of LINEMOD (reference cfg/ape.data:9-14), used as benchmark constants.
LINEMOD_INTRINSICS = dict(fx=572.4114, fy=573.5704, u0=325.2611, v0=242.0489, width=640, height=480)
