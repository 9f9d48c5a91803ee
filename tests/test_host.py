"""CPU-only tests: host logic (cfg parsing, execution plan, .weights format), the C-ABI surface, and the N>1
gradient all-reduce path with the gloo backend (world size 2).  No GPU, no compute calls into the library."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from singleshotpose_b200 import _lib, Darknet
from singleshotpose_b200.cfg import parse_cfg, layer_shapes, print_cfg
from singleshotpose_b200.cfgs import yolo_pose_cfg_text
from singleshotpose_b200.engine import build_plan

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg_parse_semantics(cfg_path, cfg_multi_path, capsys):
    b = parse_cfg(cfg_path)
    assert len(b) == 33 and b[0]["type"] == "net" and b[-1]["type"] == "region"
    assert b[1]["batch_normalize"] == "1" and b[-2]["batch_normalize"] == 0        # default for [convolutional]
    assert b[-1]["anchors"] == "" and b[-1]["classes"] == "1"
    shp = layer_shapes(b)
    assert [s[0] for s in shp].count("conv") == 23
    assert shp[29][1:3] == (1280, 1024) and shp[30][2] == 20 and shp[30][5:7] == (13, 13)
    bm = parse_cfg(cfg_multi_path)
    assert layer_shapes(bm)[30][2] == 160 and bm[-1]["num"] == "5"
    print_cfg(b)
    out = capsys.readouterr().out
    assert "   29 conv   1024  3 x 3 / 1    13 x  13 x1280   ->    13 x  13 x1024" in out
    assert "   28 route  27 24" in out


def test_execution_plan(cfg_path):
    layers = build_plan(parse_cfg(cfg_path))
    assert len(layers) == 23
    l16 = [L for L in layers if L.block_ind == 16][0]
    assert sorted(k for (_, _, k) in l16.dests) == [_lib.ROUTE_DIRECT, _lib.ROUTE_POOL]     # maxpool 17 + route 25
    l29 = [L for L in layers if L.block_ind == 29][0]
    assert l29.cin == 1280
    assert [(layers[s].block_ind, k, c0, c) for (s, k, c0, c) in l29.leaves] == [(26, _lib.ROUTE_REORG, 0, 256), (24, _lib.ROUTE_DIRECT, 256, 1024)]
    assert layers[0].first and layers[0].k_cin == 32 and layers[0].k_taps == 1
    assert not layers[-1].bn and layers[-1].cout == 20


def test_unsupported_blocks_raise(tmp_path):
    txt = yolo_pose_cfg_text().replace("[maxpool]\nsize=2\nstride=2", "[maxpool]\nsize=2\nstride=1", 1)
    p = tmp_path / "bad.cfg"
    p.write_text(txt)
    with pytest.raises(NotImplementedError):
        Darknet(str(p))


def test_parameter_names_and_counts(cfg_path):
    torch.manual_seed(0)
    m = Darknet(cfg_path)
    names = [n for n, _ in m.named_parameters()]
    assert len(names) == 68 and names[0] == "models.0.conv1.weight" and names[1] == "models.0.bn1.weight"
    assert names[-2:] == ["models.30.conv23.weight", "models.30.conv23.bias"]
    assert sum(p.numel() for p in m.parameters()) == 50547764
    assert (m.width, m.height, m.test_width, m.num_keypoints, m.num_classes, m.num_anchors) == (416, 416, 672, 9, 1, 1)
    assert m.models[-1].noobject_scale == 0.1 and m.models[-1].object_scale == 5.0     # cfg values land on the unused head


def test_weights_file_format_roundtrip(cfg_path, tmp_path):
    torch.manual_seed(3)
    m = Darknet(cfg_path)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(); mod.running_var.uniform_(0.5, 2.0)
    m.seen = 4711
    f = str(tmp_path / "a.weights")
    m.save_weights(f)
    assert os.path.getsize(f) == 16 + 4 * 50568436                      # header + fp32 stream
    raw = np.fromfile(f, dtype=np.float32, offset=16)
    bn0, conv0 = m.models[0][1], m.models[0][0]
    np.testing.assert_array_equal(raw[:32], bn0.bias.detach().numpy())            # order: bn.bias, bn.weight, mean, var, conv.weight
    np.testing.assert_array_equal(raw[64:96], bn0.running_mean.numpy())
    np.testing.assert_array_equal(raw[128:128 + 864], conv0.weight.detach().numpy().reshape(-1))   # OIHW order
    torch.manual_seed(4)
    m2 = Darknet(cfg_path)
    m2.load_weights(f)
    assert int(m2.seen) == 4711
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        if "num_batches" not in k:
            assert torch.equal(a, b), k
    torch.manual_seed(5)
    m3 = Darknet(cfg_path)
    last_before = m3.models[30][0].weight.detach().clone()
    m3.load_weights_until_last(f)
    assert torch.equal(m3.models[29][0].weight, m.models[29][0].weight)
    assert torch.equal(m3.models[30][0].weight, last_before)                       # last conv untouched (darknet.py:310)


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "ssp_b200.h")).read()
    declared = set(re.findall(r"\b(ssp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ssp_version() >= 100
    assert _lib.flat_alloc_rows(64, 416, 416) >= 64 * 417 * 417 + 417 + 2


def test_no_cpu_fallback(cfg_path):
    from singleshotpose_b200 import RegionLoss
    with pytest.raises(_lib.SspError):
        Darknet(cfg_path)(torch.zeros(1, 3, 416, 416))
    with pytest.raises(_lib.SspError):
        RegionLoss()(torch.zeros(1, 20, 13, 13), torch.zeros(1, 1050), 0)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "singleshotpose_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from singleshotpose_b200.optim import all_reduce_flat_, dp_hyperparams
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
g = torch.arange(10, dtype=torch.float32) * (r + 1)          # rank-dependent "gradient of a sum-loss"
all_reduce_flat_(g)
assert torch.equal(g, torch.arange(10, dtype=torch.float32) * 3), g
lr, wd = dp_hyperparams(0.001 * 0.1, 0.0005, per_gpu_batch=64)
assert abs(lr - 1e-4 / 128) < 1e-12 and abs(wd - 0.0005 * 128) < 1e-9
dist.barrier(); dist.destroy_process_group()
print("ok", r)
'''


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    ps = [subprocess.Popen([sys.executable, str(script), REPO, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
          for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in ps]
    assert all(p.returncode == 0 for p in ps), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_other_ranks_exit():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_optimizer_state_checkpoint_interchanges_with_torch_sgd(cfg_path):
    """FlatSGD.state_dict()/load_state_dict() speak torch.optim.SGD's format: momentum buffers are stored flat in
    the kernel's OHWI layout but exchanged as OIHW tensors, so a checkpoint moves between FlatSGD and train.py:388's optimiser."""
    import torch
    from singleshotpose_b200.darknet import Darknet
    from singleshotpose_b200.optim import FlatSGD
    torch.manual_seed(0)
    m = Darknet(cfg_path)
    m._engine.materialize(torch.device("cpu"))                       # flat buffers + views; no kernel involved
    params = list(m.parameters())
    ref = torch.optim.SGD(params, lr=1e-3, momentum=0.9, dampening=0, weight_decay=0.032)
    gen = torch.Generator().manual_seed(1)
    for p in params[:6] + params[-2:]:                               # optim.SGD keeps buffers only for parameters it stepped
        ref.state[p]["momentum_buffer"] = torch.randn(p.shape, generator=gen)
    sd = ref.state_dict()
    opt = FlatSGD(m, lr=5.0, momentum=0.0, weight_decay=0.0)
    opt.load_state_dict(sd)
    assert opt.param_groups[0] == dict(lr=1e-3, momentum=0.9, weight_decay=0.032)
    w0 = params[0]
    off, n, _ = m._engine._slices[id(w0)]
    want = sd["state"][0]["momentum_buffer"].permute(0, 2, 3, 1).reshape(-1)          # flat storage is [co][kh][kw][ci]
    assert torch.equal(opt._v[off:off + n], want)
    out = opt.state_dict()
    assert set(out["state"]) == set(range(len(params)))
    for i, p in enumerate(params):
        got = out["state"][i]["momentum_buffer"]
        exp = sd["state"][i]["momentum_buffer"] if i in sd["state"] else torch.zeros(p.shape)
        assert got.is_contiguous() and torch.equal(got, exp), i
    ref2 = torch.optim.SGD(params, lr=1.0)
    ref2.load_state_dict(out)                                        # and back into the stock optimiser
    assert ref2.param_groups[0]["momentum"] == 0.9 and torch.equal(ref2.state[params[3]]["momentum_buffer"], out["state"][3]["momentum_buffer"])
    with pytest.raises(ValueError):
        bad = {"state": {}, "param_groups": [dict(sd["param_groups"][0], nesterov=True)]}
        opt.load_state_dict(bad)


def test_checkpoint_weights_plus_optimizer_state(cfg_path, tmp_path):
    import torch
    from singleshotpose_b200.darknet import Darknet
    from singleshotpose_b200.optim import FlatSGD
    from singleshotpose_b200.checkpoint import save_checkpoint, load_checkpoint
    torch.manual_seed(3)
    a = Darknet(cfg_path); a._engine.materialize(torch.device("cpu"))
    oa = FlatSGD(a, lr=1e-4, momentum=0.9, weight_decay=0.032)
    oa._v = torch.randn(a._engine.flat_params.numel(), generator=torch.Generator().manual_seed(4))
    a.seen, a.iter = 6400, 100
    f = str(tmp_path / "ck.weights")
    save_checkpoint(a, oa, f)
    torch.manual_seed(9)
    b = Darknet(cfg_path); b._engine.materialize(torch.device("cpu"))
    ob = FlatSGD(b, lr=1.0)
    assert load_checkpoint(b, ob, f) is True
    assert (b.seen, b.iter) == (6400, 100) and ob.param_groups[0]["momentum"] == 0.9
    assert torch.equal(ob._v, oa._v)
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), n
    os.remove(f + ".optim.pt")
    with pytest.raises(FileNotFoundError):
        load_checkpoint(b, ob, f)
    assert load_checkpoint(b, ob, f, strict=False) is False
