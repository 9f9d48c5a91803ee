#!/bin/bash
# Round 2, call 6: branch-free bandT epilogue, software-pipelined l0 kernels: tests, A/Bs, launch list, ncu of the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "l0_fused or bandt or conv_gemm" > gpurun_out/r2_t_call6a.log 2>&1; tail -5 gpurun_out/r2_t_call6a.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "default                    ";
  SSP_L0_OCC=1 run "SSP_L0_OCC=1               ";
  SSP_BANDT=0 run "SSP_BANDT=0                ";
  SSP_BANDT=0 SSP_L0=direct run "SSP_BANDT=0 SSP_L0=direct  "; } | tee gpurun_out/r2_ab_call6.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 480 --csv --log-file gpurun_out/r2_launches6.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/r2_bench_under_ncu6.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches6.csv seq > gpurun_out/r2_launches6.txt 2>&1; head -30 gpurun_out/r2_launches6.txt
SSP_OVERLAP=0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"conv_bandt|l0_gram|l0_fused|l0_bwd_kernel" -c 14 -o gpurun_out/r2_new6 python tools/one_step.py 64 > gpurun_out/r2_ncu_new6.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_new6.ncu-rep > gpurun_out/r2_new6_ncu_full.txt; grep -E "^==|time_duration|stalls|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|issue_active" gpurun_out/r2_new6_ncu_full.txt | head -60
