#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "l0_fused" 2>&1 | tail -2
SSP_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"l0_" -c 9 --csv --log-file gpurun_out/r2_l0_24.csv python tools/one_step.py 64 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_l0_24.csv')) if len(r)>5]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[1:]: print("%-40s %8.1f us" % (r[k][:40], float(r[v].replace(',',''))/1e3))
PY
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f (%.2f)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
{ run "shift-correlation Gram"; SSP_L0_GRAM=brute run "SSP_L0_GRAM=brute     "; run "shift-correlation     "; SSP_L0_GRAM=brute run "SSP_L0_GRAM=brute     "; } | tee gpurun_out/r2_ab_call24.log
