#!/usr/bin/env python
"""PnP goldens beyond sigma <= 1 px (missing item 3): the reference's own `pnp` (utils.py:86-100 -> cv2.solvePnP ITERATIVE
+ cv2.Rodrigues) run unmodified on (a) sigma = 5 / 20 / 80 px keypoint noise and (b) the keypoints a random-init network emits
(decode of random logits through the reference-pinned oracle decode, i.e. garbage correspondences: valid.py:119-153 at random init).
Run in the build container (needs /root/reference and cv2); writes tests/golden/pnp_noise.npz.

    python tests/golden/make_golden_pnp_noise.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle.decode_ref import get_region_boxes_ref            # noqa: E402
from oracle.pnp_ref import pnp_ref                            # noqa: E402
from singleshotpose_b200 import synth                         # noqa: E402


def main():
    sys.path.insert(0, REF)
    cwd = os.getcwd(); os.chdir(REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import utils as ref_utils
    os.chdir(cwd)
    res = {}
    n = 64
    sets = []
    for sigma in (5.0, 20.0, 80.0):
        pr = synth.pnp_problems(n, sigma=sigma, seed=int(70 + sigma))
        sets.append(("s%d" % int(sigma), pr["uv"], pr["P3"], pr["K"]))
    # keypoints of a random-init network: per-image decode (get_region_boxes, utils.py:216-296) of random logits
    g = torch.Generator().manual_seed(23)
    out = torch.randn(n, 20, 13, 13, generator=g) * 0.7
    uv = np.zeros((n, 9, 2), np.float32)
    for i in range(n):
        box = np.array([float(v) for v in get_region_boxes_ref(out[i:i + 1], 1, 9)], np.float32)
        uv[i] = box[:18].reshape(9, 2) * np.array([640.0, 480.0], np.float32)      # valid.py:138-146 (im_width, im_height)
    sets.append(("net", uv, pr["P3"], pr["K"]))
    for tag, uvs, P3, K in sets:
        Rs, ts, worst = [], [], 0.0
        for i in range(n):
            R, t = ref_utils.pnp(P3, uvs[i], K)
            Ro, to = pnp_ref(P3, uvs[i], K)
            ang = np.degrees(np.arccos(np.clip((np.trace(R @ Ro.T) - 1) / 2, -1, 1)))
            worst = max(worst, ang, np.abs(t - to).max() * 1000)
            Rs.append(R); ts.append(t.reshape(3))
        print("pnp %s: numpy oracle vs reference pnp (cv2 %s) worst (deg|mm) %.2e" % (tag, __import__("cv2").__version__, worst))
        assert worst < 1e-3, (tag, worst)
        res["uv_" + tag] = uvs.astype(np.float32); res["R_" + tag] = np.array(Rs); res["t_" + tag] = np.array(ts)
    np.savez_compressed(os.path.join(HERE, "pnp_noise.npz"), P3=P3, K=K, **res)


if __name__ == "__main__":
    main()
