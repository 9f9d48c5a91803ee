#!/bin/bash
# Round 2, call 12: clusters of two CTA pairs with multicast operand tiles (conv_tc2, wgrad_tc2): tests, same-box A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "conv_gemm or wgrad or dgrad or tc2" > gpurun_out/r2_t_call12a.log 2>&1; tail -6 gpurun_out/r2_t_call12a.log
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/r2_t_call12b.log 2>&1; tail -4 gpurun_out/r2_t_call12b.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "default (clusters of 4)    ";
  SSP_TC2_CLUSTER=2 run "SSP_TC2_CLUSTER=2          ";
  run "default (repeat)           ";
  SSP_TC2_CLUSTER=2 run "SSP_TC2_CLUSTER=2 (repeat) "; } | tee gpurun_out/r2_ab_call12.log
