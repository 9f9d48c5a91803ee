#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_tc_precision.py > gpurun_out/diag.log 2>&1; tail -12 gpurun_out/diag.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/t_kernels.log | tail -15
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 600 > gpurun_out/t_net.log 2>&1; echo "net rc=$?"
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/t_net.log | tail -15
