#!/bin/bash
# Round 2, call 1: the whole GPU suite (experimental tests on, no -x), then the prepared A/Bs of round 1.
mkdir -p gpurun_out
SSP_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_t_all.log 2>&1
tail -15 gpurun_out/r2_t_all.log
bash tools/gpu_round2_first.sh 2>&1 | tee gpurun_out/r2_first.log
bash tools/gpu_round2_bn.sh 2>&1 | tee gpurun_out/r2_bn.log
