#!/bin/bash
mkdir -p gpurun_out
SSP_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"l0_" -c 9 --csv --log-file gpurun_out/r2_l0_23.csv python tools/one_step.py 64 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_l0_23.csv')) if len(r)>5]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
for r in rows[1:]: print("%-40s %s us" % (r[k][:40], r[v]))
PY
SSP_L0_GRAM=brute SSP_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"l0_gram" -c 1 --csv --log-file gpurun_out/r2_l0_23b.csv python tools/one_step.py 64 > /dev/null 2>&1; tail -1 gpurun_out/r2_l0_23b.csv | cut -c1-200
