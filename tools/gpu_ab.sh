#!/bin/bash
mkdir -p gpurun_out
for n in 64 128 256 9999; do
SSP_TC2_MIN_N=$n timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pnp > gpurun_out/b.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/b.json"))
print("tc2_min_n=$n: value %.1f img/s  %.2f ms/step  conv frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]), {k: round(v["ms_per_step"],2) for k,v in d["roofline"]["per_kind"].items()}, d["clocks"])
PY
done
