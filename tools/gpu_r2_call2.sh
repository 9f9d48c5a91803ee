#!/bin/bash
# Round 2, call 2: PnP device-vs-host probe (3 build variants), the whole GPU suite, A/Bs of this round's changes, launch list.
mkdir -p gpurun_out
python tools/probes/pnp_probe_data.py /tmp/uv.bin
for v in "-O3" "-O3 -fmad=false" "-O0 -G"; do
  nvcc -gencode arch=compute_100a,code=sm_100a $v -diag-suppress 1650 -Xcompiler -Wno-unused-result -o /tmp/pnp_probe tools/probes/pnp_probe.cu 2>/dev/null
  echo "== pnp_probe built with: $v"; timeout 120 /tmp/pnp_probe /tmp/uv.bin
done 2>&1 | tee gpurun_out/r2_pnp_probe.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_t_all2.log 2>&1
tail -12 gpurun_out/r2_t_all2.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
run "default                 "
SSP_POOL_REDUCE=full run "pooled reduce full-res  "
SSP_WGRAD_IMPL=tc run "wgrad 1-CTA             "
run "default (repeat)        "
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/r2_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches.csv seq > gpurun_out/r2_launches_bench_b64.txt 2>&1; head -32 gpurun_out/r2_launches_bench_b64.txt
