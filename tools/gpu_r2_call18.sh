#!/bin/bash
# Round 2, call 18: overlapping-row tensor maps for the 32-channel layer (block 2): wgrad boxes with two taps each, dense bands + 3 ring stages
# in the operand-swapped forward.  Tests, same-box A/B, per-launch times of the two kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "wgrad or bandt or conv_gemm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "default (overlapping-row maps)     ";
  SSP_WGRAD_OVL=0 SSP_BANDT_OVL=0 run "SSP_WGRAD_OVL=0 SSP_BANDT_OVL=0    ";
  run "default (repeat)                   "; } | tee gpurun_out/r2_ab_call18.log
SSP_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none --profile-from-start off -k regex:"conv_bandt|wgrad_tc_kernel" -c 30 python tools/one_step.py 64 2>&1 | grep -E "conv_bandt|wgrad_tc_kernel|duration|tensor|xbar" | paste - - - - | sed 's/  */ /g' | cut -c1-260 | tee gpurun_out/r2_ovl_launches.txt | head -24
