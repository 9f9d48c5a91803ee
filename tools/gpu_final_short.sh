#!/bin/bash
# final validation without the reference arm: every -m gpu test, smoke(), default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value %.1f img/s  %.2f ms/step  e2e %.1f  conv frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"]))
print("augment", d["augment"]); print("multi", d["multi"]); print("inference", d["inference"]); print("clocks", d["clocks"])
PY
