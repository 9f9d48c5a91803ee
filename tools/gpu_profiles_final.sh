#!/bin/bash
# final round-1 evidence: launch list of the bench command (eager launches, same kernels as the graph) + full captures
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 460 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
SSP_OVERLAP=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2|wgrad_tc|bn_bwd_apply|conv_band" -s 120 -c 24 -o /tmp/prof_final \
   python tools/one_step.py 64 > gpurun_out/ncu_full_final.log 2>&1
echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_final.ncu-rep > gpurun_out/prof_final.txt
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
ls -la /tmp/prof_final.ncu-rep gpurun_out/
