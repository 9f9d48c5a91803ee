#!/bin/bash
# Round-2 A/B of the BN-kernel occupancy knob: rebuild elementwise.cu with a register cap for 3 (4) blocks per SM and compare the
# training step on the same box.  The in-tree .so is rebuilt on the box only (nothing persists).
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1  %.1f img/s  %.2f ms/step' % (d['value'], d['ms_per_step']))"; }
run "default (no cap)      "
for units in 16 32; do           # more pixels per thread: fewer, longer-lived blocks (default 8)
  touch singleshotpose_b200/csrc/elementwise.cu
  SSP_BN_UNITS=$units python singleshotpose_b200/csrc/build.py > /dev/null 2>&1 || { echo "build failed"; exit 1; }
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "bn_" --timeout 200 2>&1 | tail -1
  run "SSP_BN_UNITS=$units    "
done
for mb in 3 4; do
  touch singleshotpose_b200/csrc/elementwise.cu
  SSP_BN_MINBLOCKS=$mb python singleshotpose_b200/csrc/build.py > /dev/null 2>&1 || { echo "build failed"; exit 1; }
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "bn_" --timeout 200 2>&1 | tail -1
  run "SSP_BN_MINBLOCKS=$mb"
done
