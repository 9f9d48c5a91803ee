"""GPU parity of the whole hot path through the reference-facing API (Darknet / RegionLoss) against the oracle
network on the CPU and the reference-generated golden logits.  Tolerances are the north star's: 1e-3 relative
(inf-norm over the logits) for the forward pass and the loss; weight gradients use the single-pass bf16
backward and are checked at 5e-2 (documented in DESIGN.md)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle.darknet_ref import RefDarknet
from oracle import region_loss_ref as RL
from singleshotpose_b200 import Darknet, RegionLoss, FlatSGD, synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _populate_eval(model):
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for bn in bns:
        bn.reset_running_stats(); bn.momentum = None
    model.train()
    with torch.no_grad():
        for s in (0, 10, 11):
            model(synth.images(2, seed=s))
    for bn in bns:
        bn.momentum = 0.1


@pytest.fixture(scope="module")
def pair(cfg_path):
    torch.manual_seed(0)
    ref = RefDarknet(cfg_path)
    torch.manual_seed(0)
    dut = Darknet(cfg_path)
    for a, b in zip(ref.state_dict().values(), dut.state_dict().values()):
        assert torch.equal(a, b)                                   # identical seeded initialisation
    return ref, dut.cuda()


def test_train_forward_backward_matches_oracle_and_golden(pair, golden_dir):
    ref, dut = pair
    g = np.load(os.path.join(golden_dir, "net_b2.npz"))
    x, tgt = synth.images(2, seed=0), synth.targets(2, seed=1)
    ref.train(); dut.train()
    out_ref = ref(x)
    out = dut(x.cuda())
    assert _rel(out.detach().cpu(), out_ref.detach()) < 1e-3
    assert _rel(out.detach().cpu(), torch.from_numpy(g["train_logits"])) < 1e-3          # reference's own output
    # running statistics after one train forward (momentum 0.1, unbiased variance)
    np.testing.assert_allclose(dut.models[0][1].running_mean.cpu().numpy(), g["running_mean0"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(dut.models[29][1].running_var.cpu().numpy(), g["running_var29"], rtol=1e-3, atol=1e-5)
    # loss + backward
    l_ref, _ = RL.region_loss_ref(out_ref, tgt, 20)
    l_ref.backward()
    crit = RegionLoss(); crit.verbose = False
    loss = crit(out, tgt, 20)
    loss.backward()
    assert float(loss) == pytest.approx(float(l_ref), rel=1e-3)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-3)
    worst = 0.0
    for (n, p), (_, q) in zip(dut.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, n
        worst = max(worst, _rel(p.grad.cpu(), q.grad))
    assert worst < 5e-2, worst
    np.testing.assert_allclose(dut.models[0][0].weight.grad.cpu().numpy(), g["first_w_grad"],
                               atol=5e-2 * np.abs(g["first_w_grad"]).max())


def test_eval_forward_matches_oracle_and_golden(pair, golden_dir):
    ref, dut = pair
    g = np.load(os.path.join(golden_dir, "net_b2.npz"))
    ref2 = copy.deepcopy(ref)
    _populate_eval(ref2)
    dut.load_state_dict(ref2.state_dict())
    ref2.eval(); dut.eval()
    x = synth.images(2, seed=0)
    with torch.no_grad():
        o_ref = ref2(x)
        o = dut(x.cuda())
    assert _rel(o.cpu(), o_ref) < 1e-3
    assert _rel(o.cpu(), torch.from_numpy(g["eval_logits"])) < 1e-3
    # batch of one, like valid.py
    with torch.no_grad():
        o1 = dut(x[:1].cuda())
    assert _rel(o1.cpu(), o_ref[:1]) < 1e-3


def test_sgd_step_matches_torch_optimizer(cfg_path):
    """FlatSGD (one fused kernel) and torch.optim.SGD on the permuted parameter views give the same update."""
    torch.manual_seed(1)
    a = Darknet(cfg_path).cuda()
    b = copy.deepcopy(a)
    x, tgt = synth.images(2, seed=3).cuda(), synth.targets(2, seed=4)
    crit = RegionLoss(); crit.verbose = False
    opt_a = FlatSGD(a, lr=1e-3, momentum=0.9, weight_decay=0.032)
    opt_b = torch.optim.SGD(b.parameters(), lr=1e-3, momentum=0.9, dampening=0, weight_decay=0.032)
    for _ in range(2):
        for m, opt in ((a, opt_a), (b, opt_b)):
            opt.zero_grad()
            crit(m(x), tgt, 20).backward()
            opt.step()
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert _rel(p.detach(), q.detach()) < 1e-4, n


def test_other_resolution_and_weights_roundtrip(cfg_path, tmp_path):
    torch.manual_seed(2)
    m = Darknet(cfg_path).cuda().eval()
    with torch.no_grad():
        o = m(synth.images(1, 352, 480, seed=5).cuda())            # multi-resolution training shapes (dataset.py:66-90)
    assert o.shape == (1, 20, 11, 15)
    f = str(tmp_path / "m.weights")
    m.seen = 1234
    m.save_weights(f)
    m2 = Darknet(cfg_path)
    m2.load_weights(f)
    assert int(m2.seen) == 1234
    for (n, p), (_, q) in zip(m.state_dict().items(), m2.state_dict().items()):
        if "num_batches" not in n:
            assert torch.equal(p.cpu(), q), n


def test_cpu_tensor_is_rejected(cfg_path):
    from singleshotpose_b200._lib import SspError
    m = Darknet(cfg_path)
    with pytest.raises(SspError):
        m(synth.images(1))
