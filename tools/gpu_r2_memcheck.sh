#!/bin/bash
# memcheck over the hand-written kernels of round 2 (same bodies as the parity tests, tools/sanitize_kernels.py).  Summaries -> gpurun_out/ -> profiles/.
mkdir -p gpurun_out
for part in l0 bandt tc2 wgrad2 wgrad misc; do
  timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_kernels.py $part > gpurun_out/r2_memcheck_${part}.log 2>&1
  echo "memcheck $part rc=$? $(grep -E 'ERROR SUMMARY|sanitize_kernels: ran' gpurun_out/r2_memcheck_${part}.log | tr '\n' ' ')"
  grep -m 12 -A3 "Invalid\|out of bounds\|misaligned" gpurun_out/r2_memcheck_${part}.log | head -40
done 2>&1 | tee gpurun_out/r2_memcheck_summary.txt
