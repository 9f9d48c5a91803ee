#!/bin/bash
# round-1 evidence: launch list of the bench command (eager launches = the kernels the graph replays) + full captures
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 460 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
SSP_OVERLAP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2|wgrad_tc|conv_band|conv0_direct|bn_bwd_apply|bn_apply" -s 100 -c 14 -o /tmp/prof_final \
   python tools/one_step.py 64 > gpurun_out/ncu_full_final.log 2>&1
echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_final.ncu-rep > gpurun_out/prof_final.txt
