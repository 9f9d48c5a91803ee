#!/bin/bash
# full GPU validation: every -m gpu test, smoke(), default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "ref rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-400
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value %.1f img/s  %.2f ms/step  e2e %.1f  conv frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("cpu", d["cpu_baseline"]); print("pnp", d["pnp"]); print("clocks", d["clocks"])
PY
