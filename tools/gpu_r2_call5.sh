#!/bin/bash
# Round 2, call 5: tests of the new kernels (fused blocks 0-1, operand-swapped narrow conv, merged-tap wgrad), A/Bs, and
# `ncu --set full` with source of the new kernels in one batch-64 step (reports come back in gpurun_out/ for local reading).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "l0_fused or bandt or wgrad or conv_gemm" > gpurun_out/r2_t_call5a.log 2>&1; tail -15 gpurun_out/r2_t_call5a.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "default (fused L0, bandT, merged wgrad)";
  SSP_BANDT=0 run "SSP_BANDT=0                            ";
  SSP_BANDT=0 SSP_WGRAD_MERGE=0 run "SSP_BANDT=0 SSP_WGRAD_MERGE=0          ";
  SSP_BANDT=0 SSP_L0=direct run "SSP_BANDT=0 SSP_L0=direct              "; } | tee gpurun_out/r2_ab_call5.log
SSP_OVERLAP=0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"conv_bandt|l0_gram|l0_fused|l0_bwd_kernel" -c 14 -o gpurun_out/r2_new python tools/one_step.py 64 > gpurun_out/r2_ncu_new.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_new.ncu-rep > gpurun_out/r2_new_ncu_full.txt; grep -E "^==|time_duration|stalls|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|issue_active" gpurun_out/r2_new_ncu_full.txt | head -80
