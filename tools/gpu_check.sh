#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -q -m gpu --timeout 600 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pnp > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value %.1f img/s  %.2f ms/step  e2e %.1f  conv frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
print(d["roofline"]["per_kind"], d["gpu_launches"], d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 450 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv | head -16
