#!/bin/bash
# ncu launch list of the bench command (cold-cache, serialised: compare shares) + one full capture of the conv kernel
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 450 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 4 -o gpurun_out/prof_conv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pnp > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/
