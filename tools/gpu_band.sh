#!/bin/bash
for bo in 0 1; do
echo "== SSP_BAND_BASEOFF=$bo"
SSP_BAND_BASEOFF=$bo timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -k "conv_gemm" 2>&1 | grep -E "passed|failed|FAILED" | head -12
done
