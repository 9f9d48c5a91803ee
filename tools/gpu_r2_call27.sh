#!/bin/bash
# Round 2, call 27: band mode of the CTA-pair kernel (3x3, <= 64 input channels, split-fp16 forward): tests, per-launch times, A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "conv_gemm or tc2" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_multi.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
for b in 1 0; do SSP_TC2_BAND=$b SSP_OVERLAP=0 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -k regex:"conv_tc2" -c 6 --csv --log-file gpurun_out/r2_tc2band_$b.csv python tools/one_step.py 64 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_tc2band_$b.csv')) if len(r)>5]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); m=h.index('Metric Name')
print("SSP_TC2_BAND=$b:", ["%s=%s" % (r[m][:12], r[v]) for r in rows[1:13]])
PY
done
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step'], d['roofline']['frac']))"; }
{ run "band mode (default)"; SSP_TC2_BAND=0 run "SSP_TC2_BAND=0     "; run "band mode          "; SSP_TC2_BAND=0 run "SSP_TC2_BAND=0     "; } | tee gpurun_out/r2_ab_call27.log
