"""Oracle: the per-image evaluation loop of the reference, valid.py:123-183, restated with numpy on the CPU.

TEST INFRASTRUCTURE ONLY (checker for utils.evaluate_poses_batched).  Every arithmetic step calls the
reference-pinned oracles: `get_region_boxes_ref` (utils.py:216-296, bit-equal to the reference, tests/golden/decode.npz) and
`pnp_ref` (utils.py:86-100 -> cv2.solvePnP, pinned to cv2 at every noise level, tests/golden/pnp*.npz); the projection /
transformation / angular-distance helpers are the three-line numpy formulas of utils.py:31-48, restated below.
Parity: pinned through those two oracles; the loop itself has no golden of its own (the reference cannot run it here: it needs
LINEMOD files and `.cuda()`), which the tests say where they use it.
"""
from __future__ import annotations

import numpy as np

from .decode_ref import get_region_boxes_ref
from .pnp_ref import pnp_ref


def calc_angular_distance(gt_rot, pr_rot):
    """utils.py:31-35"""
    trace = np.trace(np.dot(gt_rot, np.transpose(pr_rot)))
    return np.rad2deg(np.arccos((trace - 1.0) / 2.0))


def compute_projection(points_3D, transformation, internal_calibration):
    """utils.py:40-45 (fp64 math, fp32 result)"""
    out = np.zeros((2, points_3D.shape[1]), dtype="float32")
    cam = (internal_calibration.dot(transformation)).dot(points_3D)
    out[0, :] = cam[0, :] / cam[2, :]
    out[1, :] = cam[1, :] / cam[2, :]
    return out


def compute_transformation(points_3D, transformation):
    """utils.py:47-48"""
    return transformation.dot(points_3D)


def evaluate_image_ref(output_1, target_row, vertices, points_3D, K, num_classes=1, num_keypoints=9, im_width=640, im_height=480):
    """One iteration of valid.py:107-183 (batch size 1, first ground truth): output_1 (1, 2K+1+C, h, w) CPU tensor,
    target_row (>= 2K+1,) -> dict of the loop's per-sample quantities."""
    box_pr = [float(v) for v in get_region_boxes_ref(output_1, num_classes, num_keypoints)]        # valid.py:119
    n2 = 2 * num_keypoints
    box_gt = [float(target_row[j]) for j in range(1, n2 + 1)]                                       # valid.py:131-133
    c_gt = np.array(np.reshape(box_gt[:n2], [-1, 2]), dtype="float32")                              # valid.py:138-143
    c_pr = np.array(np.reshape(box_pr[:n2], [-1, 2]), dtype="float32")
    c_gt[:, 0] *= im_width; c_gt[:, 1] *= im_height
    c_pr[:, 0] *= im_width; c_pr[:, 1] *= im_height
    corner_dist = np.mean(np.linalg.norm(c_gt - c_pr, axis=1))                                      # valid.py:148-150
    P3 = np.array(points_3D, dtype="float32")
    Kf = np.array(K, dtype="float32")
    R_gt, t_gt = pnp_ref(P3, c_gt, Kf)                                                              # valid.py:152-153
    R_pr, t_pr = pnp_ref(P3, c_pr, Kf)
    trans_dist = np.sqrt(np.sum(np.square(t_gt - t_pr)))                                            # valid.py:156
    angle_dist = calc_angular_distance(R_gt, R_pr)                                                  # valid.py:160
    Rt_gt, Rt_pr = np.concatenate((R_gt, t_gt), axis=1), np.concatenate((R_pr, t_pr), axis=1)       # valid.py:164-165
    Kd = np.asarray(K, np.float64)
    p_gt, p_pr = compute_projection(vertices, Rt_gt, Kd), compute_projection(vertices, Rt_pr, Kd)   # valid.py:166-167
    pixel_dist = np.mean(np.linalg.norm(p_gt - p_pr, axis=0))                                       # valid.py:168-169
    v_gt, v_pr = compute_transformation(vertices, Rt_gt), compute_transformation(vertices, Rt_pr)   # valid.py:173-174
    vertex_dist = np.mean(np.linalg.norm(v_gt - v_pr, axis=0))                                      # valid.py:175-176
    return dict(box=np.array(box_pr, np.float32), corner_err_px=corner_dist, R_gt=R_gt, t_gt=t_gt, R_pr=R_pr, t_pr=t_pr,
                trans_err=trans_dist, angle_err_deg=angle_dist, pixel_err=pixel_dist, vertex_dist=vertex_dist)
