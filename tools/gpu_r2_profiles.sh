#!/bin/bash
# Round-2 evidence: (1) ncu launch list of the bench command (eager launches = the kernels the graph replays),
# (2) `ncu --set full` of every GEMM launch and of the HBM-bound kernels of ONE batch-64 step.  Summaries -> gpurun_out/ -> profiles/.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start on -s 1200 -c 700 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches.csv seq > gpurun_out/r2_launches_bench_b64.txt 2>&1; head -30 gpurun_out/r2_launches_bench_b64.txt
SSP_OVERLAP=0 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"conv_tc2|wgrad_tc|conv_band|conv_tc_kernel" -c 110 -o /tmp/r2_gemm python tools/one_step.py 64 > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
python tools/ncu_summary.py /tmp/r2_gemm.ncu-rep > gpurun_out/r2_gemm_ncu_full.txt
SSP_OVERLAP=0 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"bn_|conv0_direct|sgd_pack|pack_input|region_loss" -c 130 -o /tmp/r2_hbm python tools/one_step.py 64 > gpurun_out/r2_ncu_hbm.log 2>&1
echo "hbm full rc=$?"
python tools/ncu_summary.py /tmp/r2_hbm.ncu-rep > gpurun_out/r2_hbm_ncu_full.txt
python tools/step_from_full.py gpurun_out/r2_gemm_ncu_full.txt gpurun_out/r2_hbm_ncu_full.txt > gpurun_out/r2_step_table.txt 2>&1
tail -5 gpurun_out/r2_step_table.txt
