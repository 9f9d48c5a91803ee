#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -q -m gpu --timeout 600 2>&1 | grep -E "passed|failed|Error|assert" | head
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pnp > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; grep -v Warn gpurun_out/bench.err | tail -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value %.1f img/s  %.2f ms/step  e2e %.1f (%.2f ms) conv frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), d["config"]["launch_path"])
print({k: round(v["ms_per_step"],2) for k,v in d["roofline"]["per_kind"].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 450 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv 2>/dev/null | head -18
