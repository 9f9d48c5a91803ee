"""Seeded synthetic inputs for tests, smoke and bench: uniform RGB batches,
one-ground-truth label rows in the reference's 50x21 layout (dataset.py:107 / region_loss.py:29-36),
box-corner 3-D models and exact/noisy 2-D projections for PnP."""
from __future__ import annotations

import os

import numpy as np
import torch

from .cfgs import LINEMOD_INTRINSICS


def images(batch, height=416, width=416, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, height, width, generator=g)


def targets(batch, seed=1, num_keypoints=9, max_objs=50):
    """(batch, 50*21) float32: one object per image, class 0, centroid U(.1,.9), 8 corners =
    centroid + U(-.15,.15), then x/y range; remaining slots zero."""
    rng = np.random.default_rng(seed)
    nl = 2 * num_keypoints + 3
    t = np.zeros((batch, max_objs * nl), np.float32)
    for b in range(batch):
        c = rng.uniform(0.1, 0.9, size=2)
        pts = np.concatenate([c[None], c[None] + rng.uniform(-0.15, 0.15, size=(num_keypoints - 1, 2))])
        t[b, 0] = 0
        t[b, 1:1 + 2 * num_keypoints] = pts.reshape(-1)
        t[b, 1 + 2 * num_keypoints] = pts[:, 0].max() - pts[:, 0].min()
        t[b, 2 + 2 * num_keypoints] = pts[:, 1].max() - pts[:, 1].min()
    return torch.from_numpy(t)


def targets_multi(batch, seed=1, num_keypoints=9, max_objs=50, num_classes=13, min_objs=1, max_gts=3):
    """(batch, 50*21) float32 with 1..max_gts objects per image: class U{0..12}, centroid U(.1,.9),
    corners = centroid + U(-.15,.15), x/y range U(0.05, 0.4) (drives the anchor choice)."""
    rng = np.random.default_rng(seed)
    nl = 2 * num_keypoints + 3
    t = np.zeros((batch, max_objs * nl), np.float32)
    for b in range(batch):
        for k in range(int(rng.integers(min_objs, max_gts + 1))):
            c = rng.uniform(0.1, 0.9, size=2)
            pts = np.concatenate([c[None], c[None] + rng.uniform(-0.15, 0.15, size=(num_keypoints - 1, 2))])
            o = k * nl
            t[b, o] = rng.integers(0, num_classes)
            t[b, o + 1:o + 1 + 2 * num_keypoints] = pts.reshape(-1)
            t[b, o + 1 + 2 * num_keypoints] = rng.uniform(0.05, 0.4)
            t[b, o + 2 + 2 * num_keypoints] = rng.uniform(0.05, 0.4)
    return torch.from_numpy(t)


MULTI_ANCHORS = [1.4820, 2.2412, 2.0501, 3.1265, 2.3946, 4.6891, 3.1018, 3.9910, 3.4879, 5.8851]   # yolo-pose-multi.cfg:240


def intrinsics(dtype=np.float64):
    k = LINEMOD_INTRINSICS
    return np.array([[k["fx"], 0.0, k["u0"]], [0.0, k["fy"], k["v0"]], [0.0, 0.0, 1.0]], dtype)


def box_points(half_extents=(0.038, 0.039, 0.046), with_center=True):
    """(9,3) or (8,3) float32: origin + the 8 corners in get_3D_corners order (utils.py:66-84:
    x outermost, z fastest, min before max)."""
    hx, hy, hz = half_extents
    c = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    if with_center:
        c = np.concatenate([np.zeros((1, 3)), c])
    return c.astype(np.float32)


def _rodrigues(r):
    th = np.linalg.norm(r, axis=-1, keepdims=True)
    u = r / np.maximum(th, 1e-300)
    c, s = np.cos(th)[..., None], np.sin(th)[..., None]
    ux = np.zeros(r.shape[:-1] + (3, 3))
    ux[..., 0, 1], ux[..., 0, 2] = -u[..., 2], u[..., 1]
    ux[..., 1, 0], ux[..., 1, 2] = u[..., 2], -u[..., 0]
    ux[..., 2, 0], ux[..., 2, 1] = -u[..., 1], u[..., 0]
    return c * np.eye(3) + (1 - c) * u[..., :, None] * u[..., None, :] + s * ux


def pnp_problems(n, sigma=0.5, seed=5, with_center=True):
    """n synthetic PnP problems sharing one 3-D model and K.  Returns dict with P3 (N,3) f32,
    uv (n,N,2) f32, K (3,3) f32, and the generating R (n,3,3), t (n,3) in f64."""
    rng = np.random.default_rng(seed)
    P3 = box_points(with_center=with_center)
    K = intrinsics()
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    rv = ax * rng.uniform(0, np.pi, size=(n, 1))
    t = np.stack([rng.uniform(-.2, .2, n), rng.uniform(-.15, .15, n), rng.uniform(.6, 1.2, n)], 1)
    R = _rodrigues(rv)
    Pc = np.einsum("nij,kj->nki", R, P3.astype(np.float64)) + t[:, None, :]
    uv = np.stack([K[0, 0] * Pc[..., 0] / Pc[..., 2] + K[0, 2], K[1, 1] * Pc[..., 1] / Pc[..., 2] + K[1, 2]], -1)
    uv = uv + rng.normal(size=uv.shape) * sigma
    return dict(P3=P3, uv=uv.astype(np.float32), K=K.astype(np.float32), R=R, t=t)


def photo_sample(seed, ow=640, oh=480, bw=500, bh=375):
    """One synthetic training sample for the image pipeline (image.py:129-142): (image, object mask, background), uint8 HxWx3 RGB.
    Integer arithmetic only, so every platform generates the same bytes.  The mask is an ellipse (255 inside, 0 outside) with
    an anti-aliased rim of intermediate values, like the LINEMOD masks after PNG decoding."""
    rng = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:oh, 0:ow]
    base = np.stack([xx * 255 // max(ow - 1, 1), yy * 255 // max(oh - 1, 1), (xx + yy) * 255 // max(ow + oh - 2, 1)], -1)
    img = np.clip(base + rng.integers(-48, 49, (oh, ow, 3)), 0, 255).astype(np.uint8)
    cx, cy = int(rng.integers(ow // 4, 3 * ow // 4)), int(rng.integers(oh // 4, 3 * oh // 4))
    ax, ay = int(rng.integers(ow // 8, ow // 3)), int(rng.integers(oh // 8, oh // 3))
    d = ((xx - cx) * ay) ** 2 + ((yy - cy) * ax) ** 2                       # < (ax*ay)^2 inside the ellipse
    r2 = (ax * ay) ** 2
    m = np.where(d < r2 * 9 // 10, 255, np.where(d < r2, rng.integers(0, 256, (oh, ow)), 0)).astype(np.uint8)
    mask = np.repeat(m[:, :, None], 3, 2)
    by, bx = np.mgrid[0:bh, 0:bw]
    bbase = np.stack([255 - bx * 255 // max(bw - 1, 1), (bx * by) % 256, by * 255 // max(bh - 1, 1)], -1)
    bg = np.clip(bbase + rng.integers(-64, 65, (bh, bw, 3)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(img), np.ascontiguousarray(mask), np.ascontiguousarray(bg)


def label_rows(seed, n=1, num_keypoints=9):
    """n label rows [cls, x0, y0, ..., x8, y8, xrange, yrange] like the reference's labels/*.txt"""
    rng = np.random.default_rng(2000 + seed)
    rows = np.zeros((n, 2 * num_keypoints + 3))
    for r in rows:
        c = rng.uniform(0.2, 0.8, 2)
        pts = c + rng.uniform(-0.15, 0.15, (num_keypoints, 2))
        pts[0] = c
        r[0] = 0
        r[1:1 + 2 * num_keypoints] = pts.reshape(-1)
        r[-2:] = pts.max(0) - pts.min(0)
    return rows


def write_linemod_like(root, n=4, ow=160, oh=120, num_bg=3):
    """A tiny dataset tree with the reference's path conventions (image.py:130-131, train.py:309): JPEGImages/00000i.png, mask/000i.png,
    labels/00000i.txt, a background folder and the list file.  PNG throughout (lossless, so every decoder yields the same bytes).
    Returns (list file path, background file names)."""
    from PIL import Image
    base = os.path.join(root, "LINEMOD", "ape")
    for d in ("JPEGImages", "mask", "labels"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    bgdir = os.path.join(root, "VOCdevkit", "VOC2012", "JPEGImages")
    os.makedirs(bgdir, exist_ok=True)
    lines = []
    for i in range(n):
        img, mask, _bg = photo_sample(50 + i, ow, oh, 8, 8)
        name = "%06d" % i
        Image.fromarray(img).save(os.path.join(base, "JPEGImages", name + ".png"))
        Image.fromarray(mask).save(os.path.join(base, "mask", "%04d.png" % i))
        rows = label_rows(50 + i, n=1 + i % 2)
        with open(os.path.join(base, "labels", name + ".txt"), "w") as f:
            if i != 3:                                   # sample 3 has an empty label file (os.path.getsize == 0 branch)
                np.savetxt(f, rows)
        lines.append(os.path.join(base, "JPEGImages", name + ".png"))
    bgs = []
    for j in range(num_bg):
        _i, _m, bg = photo_sample(70 + j, 8, 8, 100 + 13 * j, 75 + 7 * j)
        pth = os.path.join(bgdir, "bg%d.png" % j)
        Image.fromarray(bg).save(pth)
        bgs.append(pth)
    listfile = os.path.join(root, "train.txt")
    with open(listfile, "w") as f:
        f.write("\n".join(lines) + "\n")
    return listfile, bgs
