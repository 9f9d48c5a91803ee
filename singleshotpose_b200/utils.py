"""Decode and pose utilities -- drop-in for the hot-path subset of reference utils.py.

get_region_boxes (utils.py:216-296) and pnp (utils.py:86-100) run on the GPU kernels; the small numpy helpers
(compute_projection, compute_transformation, calcAngularDistance, get_3D_corners, get_camera_intrinsic,
convert2cpu) keep the reference's names and conventions.  Batched entry points (region_boxes_batched,
pnp_batched, project_points_batched) expose the same kernels without the per-image Python loop.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr, SspError
from .utils_host import (makedirs, get_all_files, calc_pts_diameter, adi, get_2d_bb, compute_2d_bb, compute_2d_bb_from_orig_pix,  # noqa: F401
                         corner_confidences, corner_confidence, sigmoid, softmax, fix_corner_order, read_truths, read_truths_args,
                         read_pose, load_class_names, image2torch, read_data_cfg, scale_bboxes, file_lines, get_image_size, logging)


def _dev():
    if not torch.cuda.is_available():
        raise SspError("singleshotpose_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------------ host helpers
def get_camera_intrinsic(u0, v0, fx, fy):
    return np.array([[fx, 0.0, u0], [0.0, fy, v0], [0.0, 0.0, 1.0]])


def compute_projection(points_3D, transformation, internal_calibration):
    """K [R|t] X, perspective divide; float32 (2, N) like utils.py:40-45."""
    cam = internal_calibration.dot(transformation).dot(points_3D)
    out = np.zeros((2, points_3D.shape[1]), dtype="float32")
    out[0, :] = cam[0, :] / cam[2, :]
    out[1, :] = cam[1, :] / cam[2, :]
    return out


def compute_transformation(points_3D, transformation):
    return transformation.dot(points_3D)


def calcAngularDistance(gt_rot, pr_rot):
    trace = np.trace(np.dot(gt_rot, np.transpose(pr_rot)))
    return np.rad2deg(np.arccos((trace - 1.0) / 2.0))


def get_3D_corners(vertices):
    """(4, 8): min/max box corners, x outermost, z fastest (utils.py:66-84), homogeneous."""
    mn, mx = vertices[:3].min(axis=1), vertices[:3].max(axis=1)
    c = np.array([[x, y, z] for x in (mn[0], mx[0]) for y in (mn[1], mx[1]) for z in (mn[2], mx[2])])
    return np.concatenate((c.T, np.ones((1, 8))), axis=0)


def convert2cpu(gpu_matrix):
    return torch.FloatTensor(gpu_matrix.size()).copy_(gpu_matrix)


def convert2cpu_long(gpu_matrix):
    return torch.LongTensor(gpu_matrix.size()).copy_(gpu_matrix)


# ------------------------------------------------------------------------------------------ decode
def region_boxes_batched(output, num_classes, num_keypoints, only_objectness=1):
    """-> (boxes (B, 2K+3) per image, best_conf (B,), box_global (2K+3,)) as CUDA tensors."""
    if output.dim() == 3:
        output = output.unsqueeze(0)
    if not output.is_cuda:
        raise SspError("get_region_boxes runs on CUDA tensors only")
    out = output.detach().contiguous().float()
    B, C, H, W = out.shape
    assert C == 2 * num_keypoints + 1 + num_classes
    nv = 2 * num_keypoints + 3
    boxes = torch.empty(B, nv, dtype=torch.float32, device=out.device)
    best = torch.empty(B, dtype=torch.float32, device=out.device)
    glob = torch.empty(nv, dtype=torch.float32, device=out.device)
    call("ssp_region_decode_argmax", ptr(out), B, num_keypoints, num_classes, H, W, int(bool(only_objectness)),
         ptr(boxes), ptr(best), ptr(glob), stream_ptr())
    return boxes, best, glob


def get_region_boxes(output, num_classes, num_keypoints, only_objectness=1, validation=True):
    """Reference semantics: ONE box, the best cell over the whole batch -> list of 2K+3 scalars."""
    _, _, glob = region_boxes_batched(output, num_classes, num_keypoints, only_objectness)
    v = glob.cpu()
    box = [v[j] for j in range(2 * num_keypoints + 2)]
    box.append(v[2 * num_keypoints + 2].long())
    return box


# ------------------------------------------------------------------------------------------ pose
def pnp_batched(points_3D, points_2D, cameraMatrix, max_iter=20, return_iters=False):
    """points_3D (P,3) shared or (n,P,3); points_2D (n,P,2); K (3,3) -> R (n,3,3) f64, t (n,3) f64 CUDA tensors."""
    dev = _dev()
    P3 = torch.as_tensor(np.asarray(points_3D, dtype=np.float32) if not torch.is_tensor(points_3D) else points_3D)
    uv = torch.as_tensor(np.asarray(points_2D, dtype=np.float32) if not torch.is_tensor(points_2D) else points_2D)
    K = torch.as_tensor(np.asarray(cameraMatrix, dtype=np.float32) if not torch.is_tensor(cameraMatrix) else cameraMatrix)
    P3 = P3.to(dev, torch.float32).contiguous(); uv = uv.to(dev, torch.float32).contiguous(); K = K.to(dev, torch.float32).contiguous()
    if uv.dim() == 2:
        uv = uv.unsqueeze(0)
    n, npts = uv.shape[0], uv.shape[1]
    shared = P3.dim() == 2
    assert P3.shape[-2] == npts and P3.shape[-1] == 3 and uv.shape[-1] == 2
    R = torch.empty(n, 3, 3, dtype=torch.float64, device=dev)
    t = torch.empty(n, 3, dtype=torch.float64, device=dev)
    iters = torch.empty(n, dtype=torch.int32, device=dev) if return_iters else None
    call("ssp_pnp_batched", ptr(P3), 1 if shared else 0, ptr(uv), ptr(K), npts, n, max_iter, ptr(R), ptr(t), ptr(iters), stream_ptr())
    return (R, t, iters) if return_iters else (R, t)


def pnp(points_3D, points_2D, cameraMatrix):
    """Same contract as utils.py:86-100: numpy in, R (3,3) float64 and t (3,1) float64 out."""
    assert points_3D.shape[0] == points_2D.shape[0], "points 3D and points 2D must have same number of vertices"
    R, t = pnp_batched(points_3D, np.ascontiguousarray(points_2D[:, :2]), cameraMatrix)
    return R[0].cpu().numpy(), t[0].cpu().numpy().reshape(3, 1)


def project_points_batched(points_3D, Rt, internal_calibration):
    """points_3D (3|4, Nv); Rt (n,3,4) -> (n, 2, Nv) float32 CUDA tensor (compute_projection for n poses)."""
    dev = _dev()
    X = torch.as_tensor(points_3D).to(dev, torch.float32).contiguous()
    T = torch.as_tensor(Rt).to(dev, torch.float64).contiguous()
    K = torch.as_tensor(internal_calibration).to(dev, torch.float64).contiguous()
    if T.dim() == 2:
        T = T.unsqueeze(0)
    n, nv = T.shape[0], X.shape[1]
    out = torch.empty(n, 2, nv, dtype=torch.float32, device=dev)
    call("ssp_project_points", ptr(X), X.shape[0], nv, ptr(T), ptr(K), n, ptr(out), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------ batched evaluation tail
def evaluate_poses_batched(output, target, vertices, points_3D, internal_calibration, num_classes=1, num_keypoints=9,
                           im_width=640, im_height=480):
    """GPU-resident version of the per-image evaluation loop of reference valid.py:123-183: per-image decode
    (arg-max cell of EACH image, not the whole batch), PnP of the ground-truth and the predicted keypoints, reprojection of
    all mesh vertices, pixel / 3-D / angular / translation errors -- no Python loop over images.

    output (B, 2K+1+C, h, w) CUDA; target (B, >= 2K+1) rows [cls, x0, y0, ..., x8, y8, ...] (first object);
    vertices (3|4, Nv); points_3D (K, 3); internal_calibration (3, 3).  Returns a dict of CUDA tensors with leading dim B."""
    dev = output.device
    K = num_keypoints
    boxes, best, _ = region_boxes_batched(output, num_classes, K)
    B = boxes.shape[0]
    scale = torch.tensor([im_width, im_height], dtype=torch.float32, device=dev)
    pr2d = boxes[:, :2 * K].reshape(B, K, 2) * scale
    gt2d = torch.as_tensor(target)[:, 1:1 + 2 * K].to(dev, torch.float32).reshape(B, K, 2) * scale
    Kc = torch.as_tensor(np.asarray(internal_calibration, dtype=np.float32)).to(dev)
    P3 = torch.as_tensor(np.asarray(points_3D, dtype=np.float32)).to(dev)
    R, t = pnp_batched(P3, torch.cat([gt2d, pr2d], 0), Kc)             # 2B problems in one launch
    R_gt, R_pr, t_gt, t_pr = R[:B], R[B:], t[:B], t[B:]
    Rt_gt = torch.cat([R_gt, t_gt.unsqueeze(2)], 2)
    Rt_pr = torch.cat([R_pr, t_pr.unsqueeze(2)], 2)
    V = torch.as_tensor(np.asarray(vertices, dtype=np.float32)).to(dev)
    if V.shape[0] == 3:
        V = torch.cat([V, torch.ones(1, V.shape[1], device=dev)], 0)
    Kd = Kc.double()
    proj = project_points_batched(V, torch.cat([Rt_gt, Rt_pr], 0), Kd)  # (2B, 2, Nv)
    pixel_err = (proj[:B] - proj[B:]).norm(dim=1).mean(dim=1)           # valid.py:169-171 mean 2-D vertex reprojection distance
    Vd = V.double()
    tf_gt, tf_pr = Rt_gt @ Vd, Rt_pr @ Vd                               # compute_transformation
    vertex_dist = (tf_gt - tf_pr).norm(dim=1).mean(dim=1)               # valid.py:176-178
    tr = torch.einsum("bij,bij->b", R_gt, R_pr)                         # trace(R_gt R_pr^T)
    angle = torch.rad2deg(torch.arccos(((tr - 1.0) / 2.0).clamp(-1.0, 1.0)))
    return dict(boxes=boxes, conf=best, corner_err_px=(pr2d - gt2d).norm(dim=2).mean(dim=1), R_gt=R_gt, t_gt=t_gt, R_pr=R_pr, t_pr=t_pr,
                pixel_err=pixel_err, vertex_dist=vertex_dist, angle_err_deg=angle, trans_err=(t_gt - t_pr).norm(dim=1))
