#!/bin/bash
# Round 2, call 9: ncu --set full (+source) of every GEMM launch of one batch-64 step (evidence for profiles/, and the early-layer
# weight gradients / forward GEMMs that sit far below the late layers' efficiency), launch list of the bench command.
mkdir -p gpurun_out
SSP_OVERLAP=0 timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"conv_tc2|wgrad_tc|conv_bandt|conv_tc_kernel" -c 80 -o gpurun_out/r2_gemm9 python tools/one_step.py 64 > gpurun_out/r2_ncu_gemm9.log 2>&1
echo "gemm full rc=$?"; ls -la gpurun_out/r2_gemm9.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_gemm9.ncu-rep > gpurun_out/r2_gemm9_ncu_full.txt
grep -E "^==|time_duration|stalls|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_read|xbar2l1tex" gpurun_out/r2_gemm9_ncu_full.txt | head -400 > gpurun_out/r2_gemm9_short.txt; wc -l gpurun_out/r2_gemm9_short.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 480 --csv --log-file gpurun_out/r2_launches9.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/r2_bench_under_ncu9.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches9.csv seq > gpurun_out/r2_launches9.txt 2>&1; head -30 gpurun_out/r2_launches9.txt
