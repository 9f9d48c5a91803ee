"""GPU parity of the multi-object head (BASELINE.json configs[3]: yolo-pose-multi.cfg, batch 32)."""
import os

import numpy as np
import pytest
import torch

from oracle import region_loss_multi_ref as RM
from oracle.decode_multi_ref import get_multi_region_boxes_ref
from oracle.darknet_ref import RefDarknet
from singleshotpose_b200 import synth
from singleshotpose_b200.darknet_multi import Darknet
from singleshotpose_b200.region_loss_multi import RegionLoss
from singleshotpose_b200.utils_multi import get_multi_region_boxes

pytestmark = pytest.mark.gpu
A = synth.MULTI_ANCHORS


@pytest.mark.parametrize("epoch", [0, 20])
def test_region_loss_multi_matches_golden(golden_dir, epoch):
    g = np.load(os.path.join(golden_dir, "region_loss_multi.npz"))
    out = torch.from_numpy(g["output"]).cuda().requires_grad_(True)
    crit = RegionLoss(anchors=list(g["anchors"])); crit.verbose = False
    loss = crit(out, torch.from_numpy(g["target"]), epoch)
    loss.backward()
    assert float(loss) == pytest.approx(float(g["loss_e%d" % epoch]), rel=1e-4)
    ref = g["grad_e%d" % epoch]
    assert np.abs(out.grad.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    st = crit.stats()
    assert [st["nGT"], st["nCorrect"], st["nProposals"]] == list(g["counters_e%d" % epoch])
    np.testing.assert_allclose([st["loss_x"], st["loss_y"], st["loss_conf"], st["loss_cls"]], g["parts_e%d" % epoch], rtol=1e-4)


def test_region_loss_multi_matches_oracle_random():
    for seed in range(3):
        gen = torch.Generator().manual_seed(300 + seed)
        B = 4
        out = torch.randn(B, 160, 13, 13, generator=gen)
        tgt = synth.targets_multi(B, seed=400 + seed)
        # make the "previous image, last anchor" prediction match one ground truth so that tconf > 0.5 occurs
        t0 = tgt[1].numpy(); gi, gj = int(t0[1] * 13), int(t0[2] * 13)
        for k in range(9):
            vx, vy = t0[1 + 2 * k] * 13 - gi, t0[2 + 2 * k] * 13 - gj
            if k == 0:
                vx, vy = np.log(vx / (1 - vx)), np.log(vy / (1 - vy))
            out[0, 4 * 32 + 2 * k, gj, gi] = float(vx); out[0, 4 * 32 + 2 * k + 1, gj, gi] = float(vy)
        o = out.clone().requires_grad_(True)
        l_ref, info = RM.region_loss_multi_ref(o, tgt, 20, A)
        l_ref.backward()
        crit = RegionLoss(anchors=A); crit.verbose = False
        od = out.cuda().requires_grad_(True)
        l = crit(od, tgt, 20)
        l.backward()
        assert float(l) == pytest.approx(float(l_ref), rel=1e-4)
        assert (od.grad.cpu() - o.grad).abs().max() <= 1e-4 * o.grad.abs().max()
        st = crit.stats()
        assert (st["nGT"], st["nCorrect"]) == (info["nGT"], info["nCorrect"])
        if seed == 0:
            assert info["nCorrect"] >= 1


def test_get_multi_region_boxes_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_multi.npz"))
    boxes = get_multi_region_boxes(torch.from_numpy(g["output"]).cuda(), float(g["conf_thresh"]), 13, 9, list(g["anchors"]), 5,
                                   int(g["correspondingclass"]), only_objectness=0)
    assert [len(b) for b in boxes] == list(g["counts"])
    flat = np.array([bx for img in boxes for bx in img], dtype=np.float64)
    np.testing.assert_allclose(flat[:, :20], g["boxes"][:, :20], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(flat[:, 20], g["boxes"][:, 20])


def test_get_multi_region_boxes_fallback_path():
    gen = torch.Generator().manual_seed(21)
    out = torch.randn(3, 160, 13, 13, generator=gen) * 0.5
    out[:, [18 + 32 * a for a in range(5)]] -= 6.0            # objectness tiny: nothing passes the threshold -> fallback box only
    ref = get_multi_region_boxes_ref(out, 0.05, 13, 9, A, 5, 7, only_objectness=0)
    got = get_multi_region_boxes(out.cuda(), 0.05, 13, 9, A, 5, 7, only_objectness=0)
    assert [len(b) for b in got] == [len(b) for b in ref] == [1, 1, 1]
    for gb, rb in zip(got, ref):
        np.testing.assert_allclose(np.array(gb[0], dtype=np.float64), np.array([float(v) for v in rb[0]]), rtol=1e-5, atol=1e-7)


def test_multi_network_forward_b32(cfg_multi_path):
    """configs[3]: yolo-pose-multi.cfg, eval-mode forward at batch 32 against the oracle (first 2 images checked on the CPU)."""
    torch.manual_seed(0)
    ref = RefDarknet(cfg_multi_path)
    torch.manual_seed(0)
    dut = Darknet(cfg_multi_path).cuda()
    assert dut.num_anchors == 5 and dut.num_classes == 13 and isinstance(dut.models[-1], RegionLoss)
    ref.train(); dut.train()
    x = synth.images(2, seed=0)
    out_ref = ref(x)
    out = dut(x.cuda())
    assert float((out.detach().cpu() - out_ref.detach()).abs().max() / out_ref.detach().abs().max()) < 1e-3
    crit = RegionLoss(anchors=dut.anchors); crit.verbose = False
    tgt = synth.targets_multi(2, seed=5)
    loss = crit(out, tgt, 20)
    l_ref, _ = RM.region_loss_multi_ref(out_ref, tgt, 20, dut.anchors)
    assert float(loss) == pytest.approx(float(l_ref), rel=2e-3)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dut.parameters())
    dut.eval()
    with torch.no_grad():
        o32 = dut(synth.images(32, seed=3).cuda())
    assert o32.shape == (32, 160, 13, 13) and torch.isfinite(o32).all()
    boxes = get_multi_region_boxes(o32, 0.05, 13, 9, dut.anchors, 5, 3, only_objectness=0)
    assert len(boxes) == 32 and all(len(b) >= 1 for b in boxes)
