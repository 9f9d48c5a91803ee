#!/bin/bash
# Round 2, call 8: fp16 data-gradient planes (SSP_EPI_F16 / SSP_ROUTE_F16): kernel + network tests, same-box A/B.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_multi.py -m gpu -q --timeout 900 > gpurun_out/r2_t_call8.log 2>&1; tail -12 gpurun_out/r2_t_call8.log
grep -n "median over conv weights\|noise floor" gpurun_out/r2_t_call8.log | head
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms  loss %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step'], d.get('loss')))"; }
{ run "default (fp16 dX)          ";
  SSP_DX_F16=0 run "SSP_DX_F16=0               ";
  run "default (repeat)           ";
  SSP_DX_F16=0 run "SSP_DX_F16=0 (repeat)      "; } | tee gpurun_out/r2_ab_call8.log
