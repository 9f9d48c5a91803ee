#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -k "conv_gemm" 2>&1 | grep -E "passed|failed|FAILED|Error|rel" | head -20
SSP_CONV_IMPL=tc2 timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 300 2>&1 | grep -E "passed|failed|FAILED|Error" | head
for impl in tc tc2; do
SSP_CONV_IMPL=$impl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pnp > gpurun_out/bench_$impl.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$impl.json"))
print("$impl: value %.1f img/s  %.2f ms/step  conv frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]), {k: round(v["ms_per_step"],2) for k,v in d["roofline"]["per_kind"].items()})
PY
done
