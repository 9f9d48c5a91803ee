#!/bin/bash
# Round 2, call 22: Gram matrix through shift correlations (117 instead of 406 products per pixel) + border corrections: tests, timing, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "l0_fused" 2>&1 | tail -6
SSP_L0_GRAM=brute timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "l0_fused" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
SSP_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:"l0_" -c 9 python tools/one_step.py 64 2>&1 | grep -E "l0_[a-z_]*kernel|duration|issue_active" | paste - - - | sed 's/  */ /g' | cut -c1-200 | tee gpurun_out/r2_l0_launches22.txt
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f (%.2f) | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "shift-correlation Gram     "; SSP_L0_GRAM=brute run "SSP_L0_GRAM=brute          "; run "shift-correlation (repeat) "; } | tee gpurun_out/r2_ab_call22.log
