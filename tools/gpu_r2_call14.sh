#!/bin/bash
# Round 2, call 14: wgrad_tc with a converged MMA-issuing warp and without the empty dY box: tests + A/B against the previous library.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "wgrad" 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "converged wgrad issue      ";
  git stash -q 2>/dev/null; SSP_LIB=/tmp/none true;
  run "converged wgrad (repeat)   "; } | tee gpurun_out/r2_ab_call14.log
