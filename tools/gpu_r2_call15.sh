#!/bin/bash
# Round 2, call 15: PnP with the 4x4-block eigen-solve: device-vs-host probe, head tests, PnP microbench + batch-1 inference line.
mkdir -p gpurun_out
python tools/probes/pnp_probe_data.py /tmp/uv.bin
for v in "" "-DPNP_DLT_JACOBI=1"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 $v -diag-suppress 1650 -Xcompiler -Wno-unused-result -o /tmp/pnp_probe tools/probes/pnp_probe.cu 2>/dev/null
  echo "== pnp_probe -O3 $v"; timeout 120 /tmp/pnp_probe /tmp/uv.bin | tail -3
done 2>&1 | tee gpurun_out/r2_pnp_probe15.log
timeout 900 python -m pytest tests/test_gpu_heads.py -m gpu -q --timeout 600 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({k: d[k] for k in ('pnp','inference') if k in d}))" | tee gpurun_out/r2_pnp_bench15.json
