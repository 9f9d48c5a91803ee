#!/bin/bash
# Round 2, call 3: which Jacobi form survives nvcc -O3 as device code (pnp_core.h, PNP_JACOBI_VARIANT 0 / 1 / 2), then the tests that were red in call 2.
mkdir -p gpurun_out
python tools/probes/pnp_probe_data.py /tmp/uv.bin
for v in 0 1 2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DPNP_JACOBI_VARIANT=$v -diag-suppress 1650 -Xcompiler -Wno-unused-result -o /tmp/pnp_probe tools/probes/pnp_probe.cu 2>/dev/null
  echo "== pnp_probe -O3, PNP_JACOBI_VARIANT=$v"; timeout 120 /tmp/pnp_probe /tmp/uv.bin | tail -3
done 2>&1 | tee gpurun_out/r2_pnp_probe3.log
timeout 900 python -m pytest tests/test_gpu_heads.py tests/test_gpu_network.py -m gpu -q --timeout 600 -k "pnp or evaluate_poses or fused_sgd" > gpurun_out/r2_t_call3.log 2>&1
tail -15 gpurun_out/r2_t_call3.log
