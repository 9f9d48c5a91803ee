#!/bin/bash
# racecheck + synccheck (VERDICT r1 missing 5) over the hand-written kernels, CTA-pair kernel first.  Summaries -> gpurun_out/ -> profiles/.
mkdir -p gpurun_out
for tool in synccheck racecheck; do
  for part in tc2 bandt l0 wgrad2 wgrad misc; do
    timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_kernels.py $part > gpurun_out/r2_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_kernels: ran' gpurun_out/r2_${tool}_${part}.log | tr '\n' ' ')"
  done
done 2>&1 | tee gpurun_out/r2_sanitizers_summary.txt
