"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, integer/byte arithmetic) of the reference's training-image pipeline.

Follows /root/reference image.py:
  change_background   image.py:110-127   (bg.resize -> per-channel mask LUTs -> a*c + b*d -> 'L')
  data_augmentation   image.py:46-75     (jitter crop with zero fill -> resize(shape) -> random_distort_image)
  distort_image       image.py:14-32     (RGB->HSV, three 256-entry point() tables, HSV->RGB)
  rand_scale / random_distort_image  image.py:34-44
  fill_truth_detection image.py:77-108

The pixel arithmetic itself lives in the third-party Pillow library (`from PIL import Image, ImageChops, ImageMath`,
image.py:5; README.md:30 lists it unpinned; installed here: Pillow 12.2.0), so this file restates Pillow's published
algorithms:
  * Image.resize -> ImagingResample (libImaging/Resample.c): separable two-pass convolution, double-precision coefficient
    set-up (precompute_coeffs), 22-bit fixed-point 8bpc accumulation (normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
    Vertical_8bpc), 8-bit intermediate between the passes; NEAREST -> ImagingScaleAffine (Geometry.c).
  * Image.convert('HSV') / ('RGB') -> rgb2hsv_row / hsv2rgb (libImaging/Convert.c).
  * Image.point(callable) -> 256-entry table, round()ed then clipped to 8 bits (Image.py, _imaging.c getlist).
  * Image.crop outside the image -> zero fill.
Pinned: tests/test_augment_cpu.py checks every function bit-exactly against the installed Pillow (HSV both ways over every 7th of the 2^24
colours, all of them with SSP_FULL_HSV=1 -- 0 mismatches) and against the reference's own image.py functions run unmodified (goldens from tests/golden/make_golden.py).
Two Pillow-version dependences are inherited, not chosen: resize()'s default filter (BICUBIC since Pillow 7, NEAREST before)
and point()'s round() (truncation before Pillow 8.3); `resample=` selects the former explicitly.
"""
from __future__ import annotations

import math
import random as _random

import numpy as np

PRECISION_BITS = 32 - 8 - 2
NEAREST, BILINEAR, BICUBIC = 0, 2, 3            # PIL.Image.Resampling values


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x):
    if x < 0.0:
        x = -x
    if x < 1.0:
        return 1.0 - x
    return 0.0


_FILTERS = {BICUBIC: (_bicubic, 2.0), BILINEAR: (_bilinear, 1.0)}


def precompute_coeffs(in_size, in0, in1, out_size, resample):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (ksize, bounds[out,2] int, kk[out,ksize] int32)."""
    filt, fsupport = _FILTERS[resample]
    scale = float(in1 - in0) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast truncates toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(src, bounds, kk, axis):
    """one 8bpc pass along `axis` (1 = horizontal, 0 = vertical) of an HWC uint8 image"""
    src = np.moveaxis(src, axis, 0).astype(np.int64)               # resampled axis first
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for xx, (xmin, xmax) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * kk[xx, x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img, size, resample=BICUBIC):
    """PIL Image.resize(size, resample) of an HWC uint8 array; size = (width, height)."""
    img = np.ascontiguousarray(img)
    ow, oh = int(size[0]), int(size[1])
    ih, iw = img.shape[:2]
    if (iw, ih) == (ow, oh):
        return img.copy()
    if resample == NEAREST:                          # Geometry.c ImagingScaleAffine, nearest
        def tab(n_out, n_in):
            a = float(n_in) / n_out
            xo = a * 0.5
            idx = np.zeros(n_out, np.int64)
            for x in range(n_out):
                xin = int(math.floor(xo)) if xo >= 0 else -1
                idx[x] = min(max(xin, 0), n_in - 1)
                xo += a
            return idx
        return img[tab(oh, ih)][:, tab(ow, iw)]
    out = img
    _k, bh_, kh_ = precompute_coeffs(iw, 0, iw, ow, resample)
    _k, bv_, kv_ = precompute_coeffs(ih, 0, ih, oh, resample)
    if ih > iw * 100 and oh < ih:                    # Image.py resize(): very tall images are reduced vertically first (Pillow >= 11)
        return _pass(_pass(out, bv_, kv_, 0), bh_, kh_, 1) if ow != iw else _pass(out, bv_, kv_, 0)
    if ow != iw:
        out = _pass(out, bh_, kh_, 1)                # all rows: the ybox_first/last trimming in Resample.c only skips unused rows
    if oh != ih:
        out = _pass(out, bv_, kv_, 0)
    return out


def rgb2hsv_u8(rgb):
    """Convert.c rgb2hsv_row (float32 intermediates where the C code uses float, double where it promotes)."""
    rgb = np.asarray(rgb, np.uint8)
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    f32 = np.float32
    cr = (maxc - minc).astype(f32)
    safe = np.where(cr == 0, f32(1), cr)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = cr / np.where(maxc == 0, 1, maxc).astype(f32)                 # float / float -> float
        rc = (maxc - r).astype(f32) / safe
        gc = (maxc - g).astype(f32) / safe
        bc = (maxc - b).astype(f32) / safe
    # `h = 2.0 + rc - bc` is evaluated in double (2.0 is a double constant) and then stored to the float h
    h = np.where(r == maxc, (bc - gc).astype(f32),
                 np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc.astype(np.float64)).astype(f32),
                          (4.0 + gc.astype(np.float64) - rc.astype(np.float64)).astype(f32)))
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(f32)          # fmod in double, result stored to float
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    grey = minc == maxc
    out = np.stack([np.where(grey, 0, uh), np.where(grey, 0, us), maxc], -1)
    return out.astype(np.uint8)


def hsv2rgb_u8(hsv):
    """Convert.c hsv2rgb"""
    hsv = np.asarray(hsv, np.uint8)
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    f32 = np.float32
    hf = h.astype(f32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int64)
    f = (hf - i.astype(f32).astype(np.float64)).astype(f32)
    fs = (s.astype(f32).astype(np.float64) / 255.0).astype(f32)
    vd = v.astype(f32).astype(np.float64)
    fsd, fd = fs.astype(np.float64), f.astype(np.float64)
    rnd = lambda x: np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5)).astype(np.int64)     # C round(): half away from zero
    p = np.clip(rnd(vd * (1.0 - fsd)), 0, 255)
    q = np.clip(rnd(vd * (1.0 - fsd * fd)), 0, 255)
    t = np.clip(rnd(vd * (1.0 - fsd * (1.0 - fd))), 0, 255)
    sel = i % 6
    r = np.choose(sel, [v, q, p, p, t, v])
    g = np.choose(sel, [t, v, v, q, p, p])
    b = np.choose(sel, [p, p, t, v, v, q])
    grey = s == 0
    out = np.stack([np.where(grey, v, r), np.where(grey, v, g), np.where(grey, v, b)], -1)
    return out.astype(np.uint8)


def point_lut(fn):
    """Image.point(callable) on an 8-bit band: table of round(fn(i)), stored as UINT8 with clipping"""
    return np.clip(np.array([round(fn(i)) for i in range(256)], np.int64), 0, 255).astype(np.uint8)


def distort_luts(hue, sat, val):
    """the three tables of distort_image (image.py:17-27): hue shift with the reference's +-255 wrap, saturation and value gain"""
    def change_hue(x):
        x += hue * 255
        if x > 255:
            x -= 255
        if x < 0:
            x += 255
        return x
    return point_lut(change_hue), point_lut(lambda i: i * sat), point_lut(lambda i: i * val)


def distort_image(rgb, hue, sat, val):
    hsv = rgb2hsv_u8(rgb)
    lh, ls, lv = distort_luts(hue, sat, val)
    hsv = np.stack([lh[hsv[..., 0]], ls[hsv[..., 1]], lv[hsv[..., 2]]], -1)
    return hsv2rgb_u8(hsv)


def mask_luts():
    """change_background's posmask / negmask tables (image.py:121-122)"""
    return point_lut(lambda i: i / 255), point_lut(lambda i: 1 - i / 255)


def change_background(img, mask, bg, resample=BICUBIC):
    oh, ow = img.shape[:2]
    bg = resize_u8(bg, (ow, oh), resample)
    pos, neg = mask_luts()
    out = img.astype(np.int64) * pos[mask] + bg.astype(np.int64) * neg[mask]       # ImageMath works in int32
    return np.clip(out, 0, 255).astype(np.uint8)                                    # .convert('L') clips


def crop_u8(img, box):
    """Image.crop((l, t, r, b)): pixels outside the source are zero"""
    l, t, r, b = box
    h, w = img.shape[:2]
    out = np.zeros((max(b - t, 0), max(r - l, 0), img.shape[2]), np.uint8)
    x0, x1, y0, y1 = max(l, 0), min(r, w), max(t, 0), min(b, h)
    if x1 > x0 and y1 > y0:
        out[y0 - t:y1 - t, x0 - l:x1 - l] = img[y0:y1, x0:x1]
    return out


def rand_scale(s, rng=_random):
    scale = rng.uniform(1, s)
    if rng.randint(1, 10000) % 2:
        return scale
    return 1. / scale


def draw_augmentation(ow, oh, jitter, hue, saturation, exposure, rng=_random):
    """the random draws of data_augmentation + random_distort_image in the reference's order (image.py:46-75, 34-44)"""
    dw, dh = int(ow * jitter), int(oh * jitter)
    pleft, pright = rng.randint(-dw, dw), rng.randint(-dw, dw)
    ptop, pbot = rng.randint(-dh, dh), rng.randint(-dh, dh)
    flip = rng.randint(1, 10000) % 2
    dhue = rng.uniform(-hue, hue)
    dsat = rand_scale(saturation, rng)
    dexp = rand_scale(exposure, rng)
    return dict(pleft=pleft, pright=pright, ptop=ptop, pbot=pbot, flip=flip, dhue=dhue, dsat=dsat, dexp=dexp)


def data_augmentation(img, shape, jitter, hue, saturation, exposure, rng=_random, resample=BICUBIC):
    oh, ow = img.shape[:2]
    d = draw_augmentation(ow, oh, jitter, hue, saturation, exposure, rng)
    swidth, sheight = ow - d["pleft"] - d["pright"], oh - d["ptop"] - d["pbot"]
    sx, sy = float(swidth) / ow, float(sheight) / oh
    cropped = crop_u8(img, (d["pleft"], d["ptop"], d["pleft"] + swidth - 1, d["ptop"] + sheight - 1))
    dx, dy = (float(d["pleft"]) / ow) / sx, (float(d["ptop"]) / oh) / sy
    sized = resize_u8(cropped, shape, resample)
    out = distort_image(sized, d["dhue"], d["dsat"], d["dexp"])
    return out, d["flip"], dx, dy, sx, sy


def fill_truth_detection(bs, flip, dx, dy, sx, sy, num_keypoints, max_num_gt):
    """image.py:77-108 with the label rows already parsed (bs: (n, 2K+3) float64); `flip` is drawn but never applied there"""
    num_labels = 2 * num_keypoints + 3
    label = np.zeros((max_num_gt, num_labels))
    bs = np.array(bs, np.float64).reshape(-1, num_labels)
    cc = 0
    for i in range(bs.shape[0]):
        row = bs[i].copy()
        row[1] = min(0.999, max(0, row[1] * sx - dx))
        row[2] = min(0.999, max(0, row[2] * sy - dy))
        for j in range(1, num_keypoints):
            row[2 * j + 1] = row[2 * j + 1] * sx - dx
            row[2 * j + 2] = row[2 * j + 2] * sy - dy
        label[cc] = row
        cc += 1
        if cc >= 50:
            break
    return label.reshape(-1)
