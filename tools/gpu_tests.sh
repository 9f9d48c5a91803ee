#!/bin/bash
# run on the GPU box:  gpurun -- bash tools/gpu_tests.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 300 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
tail -25 gpurun_out/t_kernels.log
timeout 600 python -m pytest tests/test_gpu_heads.py -q -m gpu --timeout 300 > gpurun_out/t_heads.log 2>&1; echo "heads rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/t_heads.log
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 600 > gpurun_out/t_net.log 2>&1; echo "net rc=$?" | tee -a gpurun_out/summary.txt
tail -25 gpurun_out/t_net.log
