"""Oracle: ``cv2.solvePnP(..., flags=SOLVEPNP_ITERATIVE)`` + ``cv2.Rodrigues`` restated in numpy fp64.

The reference's ``pnp`` (utils.py:86-100) delegates all arithmetic to OpenCV
(``opencv-python``, version unpinned by the reference -- README.md:30; 4.13.0 installed
here; its source is not under /root/reference).  This file restates OpenCV 4.x's published
algorithm (``cvFindExtrinsicCameraParams2`` + ``CvLevMarq`` in calib3d):

  1. normalise the 2-D points by K (zero distortion);
  2. planarity test on the 3-D covariance (never planar for box corners + centre);
  3. DLT: L (2N x 12), smallest right singular vector of L^T L -> [RR|tt], sign by det,
     R = U V^T of svd(RR), tt scaled by ||R||_F / ||RR||_F, Rodrigues -> rvec;
  4. Levenberg-Marquardt in pixel space on (rvec, t): lambda = 10^k, k0 = -3,
     (J^T J with diag*(1+lambda)) delta = J^T e, p = p_prev - delta; on error growth k++ (<=16)
     and re-step; else k = max(k-1,-16), stop after 20 accepted iterations or when
     ||p - p_prev|| / ||p_prev|| < FLT_EPSILON.

Pinned numerically against the installed cv2 by tests/golden/make_golden.py (fixtures in
tests/golden/pnp_*.npz) and, when cv2 is importable, live in tests/test_oracle.py.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import numpy as np

FLT_EPSILON = 1.1920929e-07


def rodrigues_vec2mat(r, jac=False):
    """cv2.Rodrigues(rvec): R = cos(t) I + (1-cos t) r r^T + sin(t) [r]_x, optional dR/dr (3 x 9)."""
    r = np.asarray(r, np.float64).reshape(3)
    theta = np.linalg.norm(r)
    I = np.eye(3)
    if theta < np.finfo(np.float64).eps:
        R = I.copy()
        if not jac:
            return R
        J = np.zeros((3, 9))
        J[0, 5], J[0, 7] = -1, 1
        J[1, 2], J[1, 6] = 1, -1
        J[2, 1], J[2, 3] = -1, 1
        return R, J
    c, s = np.cos(theta), np.sin(theta)
    c1, it = 1.0 - c, 1.0 / theta
    u = r * it
    rrt = np.outer(u, u)
    rx = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    R = c * I + c1 * rrt + s * rx
    if not jac:
        return R
    drrt = np.zeros((3, 9)); drx = np.zeros((3, 9))
    for i in range(3):
        e = np.zeros(3); e[i] = 1
        drrt[i] = (np.outer(e, u) + np.outer(u, e)).reshape(9)
        drx[i] = np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]]).reshape(9)
    J = np.zeros((3, 9))
    for i in range(3):
        ri = u[i]
        a0, a1, a2 = -s * ri, (s - 2 * c1 * it) * ri, c1 * it
        a3, a4 = (c - s * it) * ri, s * it
        J[i] = a0 * I.reshape(9) + a1 * rrt.reshape(9) + a2 * drrt[i] + a3 * rx.reshape(9) + a4 * drx[i]
    return R, J


def rodrigues_mat2vec(R):
    """cv2.Rodrigues(R): rotation matrix -> axis-angle (OpenCV's branch structure)."""
    R = np.asarray(R, np.float64)
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    rv = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((rv ** 2).sum() * 0.25)
    c = np.clip((np.trace(R) - 1) * 0.5, -1.0, 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5; rx = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5; ry = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5; rz = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1, 2] > 0) != (ry * rz > 0):
            rz = -rz
        v = np.array([rx, ry, rz])
        return v * (theta / np.linalg.norm(v))
    return rv * (0.5 / s) * theta


def project(M, r, t, K, jac=False):
    """cv2.projectPoints with zero distortion.  M (N,3) -> (N,2) pixels [+ J (2N,6)]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    if jac:
        R, dRdr = rodrigues_vec2mat(r, True)
    else:
        R = rodrigues_vec2mat(r)
    P = M @ R.T + t
    z = 1.0 / P[:, 2]
    x, y = P[:, 0] * z, P[:, 1] * z
    uv = np.stack([fx * x + cx, fy * y + cy], 1)
    if not jac:
        return uv
    N = M.shape[0]
    J = np.zeros((2 * N, 6))
    for j in range(3):
        dR = dRdr[j].reshape(3, 3)
        dP = M @ dR.T
        dx = (dP[:, 0] - x * dP[:, 2]) * z
        dy = (dP[:, 1] - y * dP[:, 2]) * z
        J[0::2, j] = fx * dx
        J[1::2, j] = fy * dy
    J[0::2, 3] = fx * z; J[0::2, 5] = -fx * x * z
    J[1::2, 4] = fy * z; J[1::2, 5] = -fy * y * z
    return uv, J


def dlt_init(M, mn):
    """Non-planar initialisation of cvFindExtrinsicCameraParams2 (DLT)."""
    N = M.shape[0]
    L = np.zeros((2 * N, 12))
    for i in range(N):
        x, y = -mn[i, 0], -mn[i, 1]
        X = np.append(M[i], 1.0)
        L[2 * i, 0:4] = X;      L[2 * i, 8:12] = x * X
        L[2 * i + 1, 4:8] = X;  L[2 * i + 1, 8:12] = y * X
    LL = L.T @ L
    _, _, Vt = np.linalg.svd(LL)
    RRt = Vt[11].reshape(3, 4)
    if np.linalg.det(RRt[:, :3]) < 0:
        RRt = -RRt
    RR, tt = RRt[:, :3], RRt[:, 3]
    sc = np.linalg.norm(RR)
    U, _, Vt2 = np.linalg.svd(RR)
    R = U @ Vt2
    t = tt * (np.linalg.norm(R) / sc)
    return rodrigues_mat2vec(R), t


def solve_pnp_iterative(points_3D, points_2D, K, max_iter=20, return_info=False):
    """Returns (rvec (3,), tvec (3,)) in fp64, following OpenCV's ITERATIVE solver."""
    M = np.asarray(points_3D, np.float64).reshape(-1, 3)
    m = np.asarray(points_2D, np.float64).reshape(-1, 2)
    K = np.asarray(K, np.float64)
    mn = np.stack([(m[:, 0] - K[0, 2]) / K[0, 0], (m[:, 1] - K[1, 2]) / K[1, 1]], 1)
    Mc = M.mean(0)
    W = np.linalg.svd((M - Mc).T @ (M - Mc), compute_uv=False)
    if W[2] / W[1] < 1e-3:
        raise NotImplementedError("planar object: homography initialisation not modelled (never hit by the hot path)")
    r, t = dlt_init(M, mn)
    # --- CvLevMarq(6, 2N, max_iter, FLT_EPSILON, completeSymm) ---
    p = np.concatenate([r, t])
    lam_lg10, iters = -3, 0

    def step(JtJ, Jte, prev, lg):
        lam = np.exp(lg * np.log(10.0))
        A = JtJ.copy()
        A[np.diag_indices(6)] *= 1.0 + lam
        delta = np.linalg.lstsq(A, Jte, rcond=None)[0]
        return prev - delta

    while True:
        uv, J = project(M, p[:3], p[3:], K, jac=True)
        err = (uv - m).reshape(-1)
        JtJ, Jte = J.T @ J, J.T @ err
        prev = p.copy()
        p = step(JtJ, Jte, prev, lam_lg10)
        if iters == 0:
            prev_err = np.linalg.norm(err)
        while True:
            e = np.linalg.norm((project(M, p[:3], p[3:], K) - m).reshape(-1))
            if e > prev_err:
                lam_lg10 += 1
                if lam_lg10 <= 16:
                    p = step(JtJ, Jte, prev, lam_lg10)
                    continue
            break
        lam_lg10 = max(lam_lg10 - 1, -16)
        iters += 1
        if iters >= max_iter or np.linalg.norm(p - prev) / np.linalg.norm(prev) < FLT_EPSILON:
            break
        prev_err = e
    if return_info:
        return p[:3].copy(), p[3:].copy(), dict(iters=iters, err=e)
    return p[:3].copy(), p[3:].copy()


def pnp_ref(points_3D, points_2D, cameraMatrix):
    """Same contract as reference utils.pnp (utils.py:86-100): -> R (3,3) f64, t (3,1) f64."""
    r, t = solve_pnp_iterative(points_3D, np.asarray(points_2D)[:, :2], cameraMatrix)
    return rodrigues_vec2mat(r), t.reshape(3, 1)
