#!/bin/bash
# targeted check of the inference-mode fused epilogue: parity tests + A/B inference latency
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -x -k "eval or resolution" --timeout 300 2>&1 | tail -5
timeout 300 python - <<'PY' 2>&1 | tail -8
import torch, time, sys
sys.path.insert(0, ".")
from singleshotpose_b200.darknet import Darknet
from singleshotpose_b200 import synth, cfgs
m = Darknet(cfgs.write_cfg()).cuda().eval()
eng = m._engine
for B in (1, 64):
    x = synth.images(B, seed=1).cuda()
    for fuse in (False, True):
        eng.fuse_eval = fuse
        with torch.no_grad():
            for _ in range(5): m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): m(x)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("B=%d fuse=%s  %.3f ms  %.0f img/s" % (B, fuse, ms, B / ms * 1e3))
PY
