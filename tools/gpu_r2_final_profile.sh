#!/bin/bash
# Round 2, final evidence: per-launch ncu table of ONE batch-64 step of the final kernels (all launches, light sections) and the launch list of the bench command.
mkdir -p gpurun_out
SSP_OVERLAP=0 timeout 1500 ncu --profile-from-start off --clock-control none \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes.sum.per_second \
  -o /tmp/step_prof python tools/one_step.py 64 > gpurun_out/step_prof.log 2>&1
echo "rc=$?"
ncu -i /tmp/step_prof.ncu-rep --page raw --csv > /tmp/step_raw.csv 2>/dev/null
python tools/step_table.py /tmp/step_raw.csv > gpurun_out/r2_step_b64_per_launch.txt 2>&1; tail -3 gpurun_out/r2_step_b64_per_launch.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 460 --csv --log-file gpurun_out/r2_launches_final.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/r2_bench_under_ncu_final.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_final.csv seq > gpurun_out/r2_launches_bench_b64_final.txt 2>&1; head -34 gpurun_out/r2_launches_bench_b64_final.txt
