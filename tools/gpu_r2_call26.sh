#!/bin/bash
# Round 2, call 26: work-per-thread sweep of the BN kernels on top of the 3-blocks/SM cap (build-time knobs), same box.
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1  %.1f img/s  %.2f ms/step' % (d['value'], d['ms_per_step']))"; }
cp singleshotpose_b200/csrc/libssp_b200.so /tmp/lib_default.so
{ run "default (units 8, reduce 32)  ";
  for cfg in "16 32" "32 32" "8 16" "8 64" "4 32"; do
    set -- $cfg
    SSP_BN_UNITS=$1 SSP_BN_REDUCE_UNITS=$2 python singleshotpose_b200/csrc/build.py --force > /dev/null 2>&1
    run "units $1, reduce units $2        "
  done
  cp /tmp/lib_default.so singleshotpose_b200/csrc/libssp_b200.so
  run "default (repeat)              "; } | tee gpurun_out/r2_ab_call26.log
