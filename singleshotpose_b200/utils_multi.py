"""Multi-object decode -- drop-in for ``get_multi_region_boxes`` of reference multi_obj_pose_estimation/utils_multi.py:266-382.
The dense per-(cell, anchor) arithmetic and the reference's sequential fallback maxima run on the GPU; the host only
applies the confidence mask (one device->host copy).  The pose helpers are shared with utils.py."""
from __future__ import annotations

import torch

from ._lib import call, ptr, stream_ptr, SspError
from .utils import (pnp, pnp_batched, compute_projection, compute_transformation, calcAngularDistance, get_3D_corners,  # noqa: F401
                    get_camera_intrinsic, convert2cpu, convert2cpu_long, project_points_batched)
from .utils_host import (makedirs, get_all_files, calc_pts_diameter, adi, get_2d_bb, corner_confidences, corner_confidence,  # noqa: F401
                         sigmoid, softmax, read_truths, read_truths_args, read_pose, load_class_names, image2torch, scale_bboxes,
                         file_lines, get_image_size, logging)
from . import utils_host as _host


def read_data_cfg(datacfg):
    """utils_multi.py:428-443: as utils.read_data_cfg, but 'gpus' defaults to '0,1,2,3'"""
    options = _host.read_data_cfg(datacfg)
    with open(datacfg, 'r') as fp:
        if not any(line.split('=')[0].strip() == 'gpus' for line in fp if '=' in line):
            options['gpus'] = '0,1,2,3'
    return options


def bbox_iou(box1, box2, x1y1x2y2=False):
    """utils_multi.py:125-156: IoU of two boxes given as corners (x1y1x2y2) or as centre + size; 0.0 when they do not overlap"""
    if x1y1x2y2:
        l1, t1, r1, b1 = box1[0], box1[1], box1[2], box1[3]
        l2, t2, r2, b2 = box2[0], box2[1], box2[2], box2[3]
        w1, h1, w2, h2 = r1 - l1, b1 - t1, r2 - l2, b2 - t2
    else:
        w1, h1, w2, h2 = box1[2], box1[3], box2[2], box2[3]
        l1, r1, t1, b1 = box1[0] - w1 / 2.0, box1[0] + w1 / 2.0, box1[1] - h1 / 2.0, box1[1] + h1 / 2.0
        l2, r2, t2, b2 = box2[0] - w2 / 2.0, box2[0] + w2 / 2.0, box2[1] - h2 / 2.0, box2[1] + h2 / 2.0
    cw = w1 + w2 - (max(r1, r2) - min(l1, l2))          # overlap = sum of the sizes minus the extent of the union box
    ch = h1 + h2 - (max(b1, b2) - min(t1, t2))
    if cw <= 0 or ch <= 0:
        return 0.0
    carea = cw * ch
    return carea / (w1 * h1 + w2 * h2 - carea)


def nms(boxes, nms_thresh):
    """utils_multi.py:223-241: greedy suppression in decreasing box[4] order; suppressed boxes get box[4] = 0 IN PLACE (like the
    reference) and are left out of the returned list"""
    if len(boxes) == 0:
        return boxes
    keys = torch.zeros(len(boxes))
    for i in range(len(boxes)):
        keys[i] = 1 - boxes[i][4]
    _, order = torch.sort(keys)
    kept = []
    for i in range(len(boxes)):
        bi = boxes[order[i]]
        if bi[4] > 0:
            kept.append(bi)
            for j in range(i + 1, len(boxes)):
                bj = boxes[order[j]]
                if bbox_iou(bi, bj, x1y1x2y2=False) > nms_thresh:
                    bj[4] = 0
    return kept


def fix_corner_order(corners2D_gt):
    """utils_multi.py:244-255"""
    import numpy as np
    out = np.zeros((9, 2), dtype="float32")
    for dst, src in enumerate((0, 1, 3, 5, 7, 2, 4, 6, 8)):
        out[dst, :] = corners2D_gt[src, :]
    return out


def multi_region_dense(output, num_classes, num_keypoints, num_anchors, correspondingclass, only_objectness=1):
    """-> dict of CUDA tensors in the reference's visiting order (cell-major, anchor fastest):
    boxes (B, HW*A, 2K+3), conf (B, HW*A), max_ind (B,), max_conf (B,), max_cls (B,)."""
    if output.dim() == 3:
        output = output.unsqueeze(0)
    if not output.is_cuda:
        raise SspError("get_multi_region_boxes runs on CUDA tensors only")
    out = output.detach().contiguous().float()
    B, C, H, W = out.shape
    K, nC, nA = num_keypoints, num_classes, num_anchors
    assert C == (2 * K + 1 + nC) * nA
    n = H * W * nA
    dev = out.device
    boxes = torch.empty(B, n, 2 * K + 3, dtype=torch.float32, device=dev)
    conf = torch.empty(B, n, dtype=torch.float32, device=dev)
    det = torch.empty(B, n, dtype=torch.float32, device=dev)
    clsc = torch.empty(B, n, dtype=torch.float32, device=dev)
    max_ind = torch.empty(B, dtype=torch.int64, device=dev)
    max_conf = torch.empty(B, dtype=torch.float32, device=dev)
    max_cls = torch.empty(B, dtype=torch.float32, device=dev)
    call("ssp_region_decode_multi", ptr(out), B, K, nC, nA, H, W, int(bool(only_objectness)), int(correspondingclass), ptr(boxes),
         ptr(conf), ptr(det), ptr(clsc), ptr(max_ind), ptr(max_conf), ptr(max_cls), stream_ptr())
    return dict(boxes=boxes, conf=conf, max_ind=max_ind, max_conf=max_conf, max_cls=max_cls)


def get_multi_region_boxes(output, conf_thresh, num_classes, num_keypoints, anchors, num_anchors, correspondingclass,
                           only_objectness=1, validation=False):
    """Reference contract: list (per image) of lists of boxes [x0/w, y0/h, ..., det_conf, cls_max_conf, cls_max_id]."""
    if validation and not only_objectness:
        raise NotImplementedError("validation=True appends per-class extras; valid_multi.py does not use it")
    d = multi_region_dense(output, num_classes, num_keypoints, num_anchors, correspondingclass, only_objectness)
    K = num_keypoints
    boxes, conf = d["boxes"].cpu(), d["conf"].cpu()
    max_ind, max_conf, max_cls = d["max_ind"].cpu(), d["max_conf"].cpu(), d["max_cls"].cpu()
    flat = boxes.view(-1, 2 * K + 3)
    all_boxes = []
    for b in range(boxes.size(0)):
        sel = boxes[b][conf[b] > conf_thresh]
        cur = [[float(v) for v in row[:2 * K + 2]] + [int(row[2 * K + 2])] for row in sel]
        if len(cur) == 0 or correspondingclass not in [bx[2 * K + 2] for bx in cur]:
            src = flat[int(max_ind[b])]
            cur.append([float(v) for v in src[:2 * K]] + [float(max_conf[b]), float(max_cls[b]), int(correspondingclass)])
        all_boxes.append(cur)
    return all_boxes
