#!/bin/bash
mkdir -p gpurun_out
timeout 2400 ncu --profile-from-start off --clock-control none \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis \
  -o /tmp/step_prof python tools/one_step.py 64 > gpurun_out/step_prof.log 2>&1
echo "rc=$?"; ls -la /tmp/step_prof.ncu-rep
ncu -i /tmp/step_prof.ncu-rep --page raw --csv > gpurun_out/step_raw.csv 2>/dev/null; ls -la gpurun_out/step_raw.csv
