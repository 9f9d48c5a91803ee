#!/bin/bash
# evidence for the image pipeline: per-launch times + DRAM bytes (ncu), memcheck
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"aug_" --csv \
   --log-file gpurun_out/aug_launches.csv python tools/aug_batch.py 4 > gpurun_out/aug_under_ncu.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/aug_under_ncu.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/aug_batch.py 2 > gpurun_out/aug_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/aug_memcheck.log
