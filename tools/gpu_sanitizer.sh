#!/bin/bash
# compute-sanitizer memcheck over the kernel tests (small shapes); summary goes to profiles/
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 -x -k "conv_gemm_matches_torch or wgrad or bn_apply or conv0 or pack_layout" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log | tail -3
timeout 300 python -m pytest tests/test_gpu_heads.py -q -m gpu -k "without_ground_truth or minimum_points" 2>&1 | grep -E "passed|failed|Error" | head -3
