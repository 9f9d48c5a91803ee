#!/bin/bash
mkdir -p gpurun_out
for impl in tc tc2; do
SSP_CONV_IMPL=$impl timeout 900 ncu --set full --clock-control none -k regex:conv_tc -s 148 -c 3 -o /tmp/prof_$impl python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pnp > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_$impl.ncu-rep > gpurun_out/prof_$impl.txt
ncu -i /tmp/prof_$impl.ncu-rep --page raw --csv > gpurun_out/prof_${impl}_raw.csv 2>/dev/null
done
head -20 gpurun_out/prof_tc.txt; head -20 gpurun_out/prof_tc2.txt
