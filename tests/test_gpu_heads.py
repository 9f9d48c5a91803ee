"""GPU parity of the loss head, decode and PnP kernels against the oracle and the reference-generated goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import region_loss_ref as RL
from oracle.decode_ref import get_region_boxes_ref
from oracle.pnp_ref import pnp_ref
from singleshotpose_b200 import RegionLoss, synth, utils

pytestmark = pytest.mark.gpu


def _ang(Ra, Rb):
    return np.degrees(np.arccos(np.clip((np.trace(Ra @ Rb.T) - 1) / 2, -1, 1)))


@pytest.mark.parametrize("epoch", [0, 20])
def test_region_loss_matches_golden(golden_dir, epoch, capsys):
    g = np.load(os.path.join(golden_dir, "region_loss.npz"))
    out = torch.from_numpy(g["output"]).cuda().requires_grad_(True)
    tgt = torch.from_numpy(g["target"])                       # stays on the CPU like train.py:82-97
    crit = RegionLoss()
    loss = crit(out, tgt, epoch)
    loss.backward()
    assert float(loss) == pytest.approx(float(g["loss_e%d" % epoch]), rel=1e-4)          # tolerance: 1e-3 (north star)
    ref = g["grad_e%d" % epoch]
    assert np.abs(out.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    st = crit.stats()
    assert [st["nGT"], st["nCorrect"], st["nProposals"]] == list(g["counters_e%d" % epoch])
    np.testing.assert_allclose([st["loss_x"], st["loss_y"], st["loss_conf"]], g["parts_e%d" % epoch], rtol=1e-4)
    assert "nGT 6, recall 1, proposals 954" in capsys.readouterr().out                    # the reference's log line


def test_region_loss_matches_oracle_random():
    for seed in range(3):
        gen = torch.Generator().manual_seed(100 + seed)
        B = 5
        out = torch.randn(B, 20, 13, 13, generator=gen)
        tgt = synth.targets(B, seed=200 + seed)
        o = out.clone().requires_grad_(True)
        l_ref, info = RL.region_loss_ref(o, tgt, 20)
        l_ref.backward()
        crit = RegionLoss(); crit.verbose = False
        od = out.cuda().requires_grad_(True)
        l = crit(od, tgt.double(), 20)                         # float64 targets as from the train loader
        l.backward()
        assert float(l) == pytest.approx(float(l_ref), rel=1e-4)
        assert (od.grad.cpu() - o.grad).abs().max() <= 1e-4 * o.grad.abs().max()


def test_get_region_boxes_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode.npz"))
    box = utils.get_region_boxes(torch.from_numpy(g["output"]).cuda(), 1, 9)
    got = np.array([float(v) for v in box])
    np.testing.assert_allclose(got, g["box"], rtol=1e-5, atol=1e-6)
    assert int(box[20]) == int(g["box"][20])


def test_get_region_boxes_per_image_and_oracle():
    gen = torch.Generator().manual_seed(9)
    out = torch.randn(4, 20, 13, 13, generator=gen)
    boxes, best, glob = utils.region_boxes_batched(out.cuda(), 1, 9)
    for b in range(4):
        ref = np.array([float(v) for v in get_region_boxes_ref(out[b:b + 1], 1, 9)])
        np.testing.assert_allclose(boxes[b].cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    ref = np.array([float(v) for v in get_region_boxes_ref(out, 1, 9)])
    np.testing.assert_allclose(glob.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["s0", "s1"])
def test_pnp_matches_cv2_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "pnp.npz"))
    R, t = utils.pnp_batched(g["P3"], g["uv_" + tag], g["K"])
    R, t = R.cpu().numpy(), t.cpu().numpy()
    for i in range(64):
        assert _ang(R[i], g["R_" + tag][i]) < 1e-2                      # degrees  (north star: 1e-2 deg)
        assert np.abs(t[i] - g["t_" + tag][i]).max() * 1e3 < 1e-2       # mm       (north star: 1e-2 mm)


@pytest.mark.parametrize("tag", ["s5", "s20", "s80", "net"])
def test_pnp_matches_cv2_golden_noisy_and_network_keypoints(golden_dir, tag):
    """beyond sigma <= 1 px: 5 / 20 / 80 px noise and the garbage keypoints a random-init network emits (per-image
    decode of random logits) -- the reference's pnp (cv2.solvePnP ITERATIVE) reaches these answers only through OpenCV's exact
    DLT initialisation + LM schedule, which the kernel restates (12x12 normal matrix on raw coordinates, CvLevMarq lambda schedule)"""
    g = np.load(os.path.join(golden_dir, "pnp_noise.npz"))
    R, t = utils.pnp_batched(g["P3"], g["uv_" + tag], g["K"])
    R, t = R.cpu().numpy(), t.cpu().numpy()
    ang = np.array([_ang(R[i], g["R_" + tag][i]) for i in range(64)])
    dt = np.abs(t - g["t_" + tag]).max(axis=1) * 1e3
    assert ang.max() < 1e-2, (tag, ang.max(), int((ang >= 1e-2).sum()))     # degrees (north star: 1e-2 deg)
    assert dt.max() < 1e-2, (tag, dt.max())                                   # mm      (north star: 1e-2 mm)


def test_pnp_work_counters():
    """ssp_pnp_batched_work reports the DLT eigen-solve's work (<= 0: -(Rayleigh-quotient steps) of the 4x4-block solve, > 0: sweeps of the
    12x12 Jacobi fall-back) / LM iterations / LM solves per problem (what bench.py's FLOP/s figure uses)"""
    from singleshotpose_b200._lib import call, ptr, stream_ptr
    pr = synth.pnp_problems(256, sigma=0.5, seed=5)
    P3 = torch.from_numpy(pr["P3"]).cuda(); uv = torch.from_numpy(pr["uv"]).cuda(); K = torch.from_numpy(pr["K"]).cuda()
    R = torch.empty(256, 9, dtype=torch.float64, device="cuda"); t = torch.empty(256, 3, dtype=torch.float64, device="cuda")
    work = torch.zeros(256, 3, dtype=torch.int32, device="cuda")
    call("ssp_pnp_batched_work", ptr(P3), 1, ptr(uv), ptr(K), 9, 256, 20, ptr(R), ptr(t), ptr(work), stream_ptr())
    w = work.cpu().numpy()
    assert (w[:, 0] >= -24).all() and (w[:, 0] <= 30).all()
    assert (w[:, 0] <= 0).mean() > 0.95                              # clean 0.5 px problems: the block solve, not the fall-back
    assert (w[:, 1] >= 1).all() and (w[:, 1] <= 20).all() and (w[:, 2] >= w[:, 1]).all()
    R2, t2 = utils.pnp_batched(pr["P3"], pr["uv"], pr["K"])
    assert torch.equal(R2.reshape(256, 9), R) and torch.equal(t2.reshape(256, 3), t)


def test_pnp_reference_signature_and_8_points():
    pr = synth.pnp_problems(4, sigma=0.5, seed=3, with_center=False)    # 8-point variant
    for i in range(4):
        R, t = utils.pnp(pr["P3"], pr["uv"][i], pr["K"])
        assert R.shape == (3, 3) and t.shape == (3, 1) and R.dtype == np.float64
        Ro, to = pnp_ref(pr["P3"], pr["uv"][i], pr["K"])
        assert _ang(R, Ro) < 1e-2 and np.abs(t - to).max() * 1e3 < 1e-2


def test_pnp_large_batch_properties():
    """1e5 problems: exact data => the generating pose is recovered; reprojection error is tiny."""
    pr = synth.pnp_problems(100000, sigma=0.0, seed=5)
    R, t, iters = utils.pnp_batched(pr["P3"], pr["uv"], pr["K"], return_iters=True)
    R, t = R.cpu().numpy(), t.cpu().numpy()
    tr = np.clip((np.einsum("nij,nij->n", R, pr["R"]) - 1) / 2, -1, 1)
    ang = np.degrees(np.arccos(tr))
    assert np.percentile(ang, 99.9) < 1e-2 and np.abs(t - pr["t"]).max() * 1e3 < 0.05    # float32 inputs limit exactness
    assert int(iters.max()) <= 20


def test_project_points_matches_numpy():
    pr = synth.pnp_problems(3, sigma=0.0, seed=1)
    X = np.concatenate([pr["P3"].T.astype(np.float64), np.ones((1, 9))])
    Rt = np.concatenate([pr["R"], pr["t"][:, :, None]], axis=2)
    K = synth.intrinsics()
    got = utils.project_points_batched(X.astype(np.float32), Rt, K).cpu().numpy()
    for i in range(3):
        ref = utils.compute_projection(X, Rt[i], K)
        np.testing.assert_allclose(got[i], ref, rtol=1e-6, atol=1e-4)


def test_evaluate_poses_batched_matches_per_image_loop():
    """8(f).1: the batched evaluation tail == the reference's per-image loop (valid.py:123-183) built from the
    reference-API functions (get_region_boxes / pnp / compute_projection / calcAngularDistance)."""
    gen = torch.Generator().manual_seed(31)
    B = 6
    pr = synth.pnp_problems(B, sigma=0.0, seed=12)
    out = torch.randn(B, 20, 13, 13, generator=gen) * 0.3
    tgt = torch.zeros(B, 21)
    for b in range(B):                                   # plant the (noisy) true keypoints in one confident cell per image
        uvn = pr["uv"][b] / np.array([640.0, 480.0], np.float32) + np.random.default_rng(b).normal(size=(9, 2)).astype(np.float32) * 1e-3
        cx, cy = int(uvn[0, 0] * 13), int(uvn[0, 1] * 13)
        cx, cy = min(max(cx, 0), 12), min(max(cy, 0), 12)
        for k in range(9):
            vx, vy = uvn[k, 0] * 13 - cx, uvn[k, 1] * 13 - cy
            if k == 0:
                vx, vy = np.log(np.clip(vx, 1e-3, 1 - 1e-3) / (1 - np.clip(vx, 1e-3, 1 - 1e-3))), np.log(np.clip(vy, 1e-3, 1 - 1e-3) / (1 - np.clip(vy, 1e-3, 1 - 1e-3)))
            out[b, 2 * k, cy, cx] = float(vx); out[b, 2 * k + 1, cy, cx] = float(vy)
        out[b, 18, cy, cx] = 6.0
        tgt[b, 1:19] = torch.from_numpy((pr["uv"][b] / np.array([640.0, 480.0], np.float32)).reshape(-1))
    rng = np.random.default_rng(3)
    verts = np.concatenate([rng.uniform(-0.04, 0.04, size=(3, 500)), np.ones((1, 500))]).astype(np.float64)
    Kc = synth.intrinsics()
    res = utils.evaluate_poses_batched(out.cuda(), tgt, verts, pr["P3"], Kc)
    for b in range(B):
        box = utils.get_region_boxes(out[b:b + 1].cuda(), 1, 9)
        c_pr = np.array([float(v) for v in box[:18]], np.float32).reshape(9, 2) * np.array([640, 480], np.float32)
        c_gt = tgt[b, 1:19].numpy().reshape(9, 2) * np.array([640, 480], np.float32)
        R_gt, t_gt = utils.pnp(pr["P3"], c_gt, Kc.astype(np.float32))
        R_pr, t_pr = utils.pnp(pr["P3"], c_pr, Kc.astype(np.float32))
        Rt_gt, Rt_pr = np.concatenate((R_gt, t_gt), 1), np.concatenate((R_pr, t_pr), 1)
        p_gt, p_pr = utils.compute_projection(verts, Rt_gt, Kc), utils.compute_projection(verts, Rt_pr, Kc)
        assert float(res["pixel_err"][b]) == pytest.approx(np.mean(np.linalg.norm(p_gt - p_pr, axis=0)), rel=1e-3, abs=1e-3)
        v_gt, v_pr = utils.compute_transformation(verts, Rt_gt), utils.compute_transformation(verts, Rt_pr)
        assert float(res["vertex_dist"][b]) == pytest.approx(np.mean(np.linalg.norm(v_gt - v_pr, axis=0)), rel=1e-3, abs=1e-6)
        assert float(res["angle_err_deg"][b]) == pytest.approx(utils.calcAngularDistance(R_gt, R_pr), abs=1e-3)
        assert float(res["trans_err"][b]) == pytest.approx(np.linalg.norm(t_gt - t_pr), rel=1e-3, abs=1e-6)


@pytest.mark.parametrize("noise", [1e-3, 3e-2])
def test_evaluate_poses_batched_matches_oracle_loop(noise):
    """8(f).1 against the ORACLE: decode_ref + pnp_ref + numpy of oracle/eval_ref.py, one valid.py:107-183 iteration per image
    (the reference evaluates with batch size 1).  noise = keypoint perturbation in normalised image units (3e-2 ~ 19 px: the
    predicted pose is far from the ground truth, PnP starts from a poor DLT)."""
    from oracle.eval_ref import evaluate_image_ref
    gen = torch.Generator().manual_seed(41)
    B = 6
    pr = synth.pnp_problems(B, sigma=0.0, seed=13)
    out = torch.randn(B, 20, 13, 13, generator=gen) * 0.3
    tgt = torch.zeros(B, 21)
    for b in range(B):
        uvn = pr["uv"][b] / np.array([640.0, 480.0], np.float32) + np.random.default_rng(b).normal(size=(9, 2)).astype(np.float32) * noise
        cx, cy = min(max(int(uvn[0, 0] * 13), 0), 12), min(max(int(uvn[0, 1] * 13), 0), 12)
        for k in range(9):
            vx, vy = uvn[k, 0] * 13 - cx, uvn[k, 1] * 13 - cy
            if k == 0:
                vx, vy = (np.log(np.clip(v, 1e-3, 1 - 1e-3) / (1 - np.clip(v, 1e-3, 1 - 1e-3))) for v in (vx, vy))
            out[b, 2 * k, cy, cx] = float(vx); out[b, 2 * k + 1, cy, cx] = float(vy)
        out[b, 18, cy, cx] = 6.0
        tgt[b, 1:19] = torch.from_numpy((pr["uv"][b] / np.array([640.0, 480.0], np.float32)).reshape(-1))
    verts = np.concatenate([np.random.default_rng(3).uniform(-0.04, 0.04, size=(3, 500)), np.ones((1, 500))]).astype(np.float64)
    Kc = synth.intrinsics()
    res = utils.evaluate_poses_batched(out.cuda(), tgt, verts, pr["P3"], Kc)
    for b in range(B):
        ref = evaluate_image_ref(out[b:b + 1], tgt[b].numpy(), verts, pr["P3"], Kc)
        np.testing.assert_allclose(res["boxes"][b].cpu().numpy(), ref["box"], rtol=1e-5, atol=1e-6)
        assert _ang(res["R_pr"][b].cpu().numpy(), ref["R_pr"]) < 1e-2 and _ang(res["R_gt"][b].cpu().numpy(), ref["R_gt"]) < 1e-2
        assert np.abs(res["t_pr"][b].cpu().numpy() - ref["t_pr"].reshape(3)).max() * 1e3 < 1e-2
        assert float(res["corner_err_px"][b]) == pytest.approx(ref["corner_err_px"], rel=1e-4, abs=1e-3)
        assert float(res["pixel_err"][b]) == pytest.approx(ref["pixel_err"], rel=1e-3, abs=1e-3)
        assert float(res["vertex_dist"][b]) == pytest.approx(ref["vertex_dist"], rel=1e-3, abs=1e-6)
        assert float(res["angle_err_deg"][b]) == pytest.approx(ref["angle_err_deg"], abs=2e-3)
        assert float(res["trans_err"][b]) == pytest.approx(ref["trans_err"], rel=1e-3, abs=1e-6)


def test_region_loss_image_without_ground_truth():
    """Edge case: an image whose label row is empty (x0 == 0).  The reference raises IndexError there
    (region_loss.py:40); here it contributes only the no-object confidence term."""
    gen = torch.Generator().manual_seed(41)
    out = torch.randn(3, 20, 13, 13, generator=gen)
    tgt = synth.targets(3, seed=42)
    tgt[1] = 0
    crit = RegionLoss(); crit.verbose = False
    od = out.cuda().requires_grad_(True)
    loss = crit(od, tgt, 20)
    loss.backward()
    st = crit.stats()
    assert st["nGT"] == 2 and torch.isfinite(loss)
    o2 = torch.cat([out[0:1], out[2:3]]).requires_grad_(True)
    l_ref, _ = RL.region_loss_ref(o2, torch.cat([tgt[0:1], tgt[2:3]]), 20)
    conf1 = torch.sigmoid(out[1, 18])
    expect = float(l_ref) + 0.5 * float((conf1 ** 2).sum())          # noobject_scale 1, tconf 0
    assert float(loss) == pytest.approx(expect, rel=1e-4)
    assert float(od.grad[1, :18].abs().max()) == 0.0


def test_pnp_minimum_points_and_argument_errors():
    from singleshotpose_b200._lib import SspError
    pr = synth.pnp_problems(3, sigma=0.0, seed=8)
    R, t = utils.pnp_batched(pr["P3"][:6], pr["uv"][:, :6], pr["K"])           # 6 points: the DLT minimum
    for i in range(3):
        assert _ang(R[i].cpu().numpy(), pr["R"][i]) < 1e-2
    with pytest.raises(SspError):
        utils.pnp_batched(pr["P3"][:5], pr["uv"][:, :5], pr["K"])              # under-determined: rejected like cv2's CV_Assert
