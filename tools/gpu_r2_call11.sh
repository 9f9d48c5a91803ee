#!/bin/bash
# Round 2, call 11: multicast fill-rate probe #2 (decides whether operand multicast across CTA pairs is worth building), quick bench.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I singleshotpose_b200/csrc -o /tmp/mc_probe2 tools/probes/mc_probe2.cu 2>/dev/null
timeout 120 /tmp/mc_probe2 2>&1 | tee gpurun_out/r2_mc_probe2.log
timeout 120 /tmp/mc_probe2 2>&1 | tee -a gpurun_out/r2_mc_probe2.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
run "default" | tee gpurun_out/r2_ab_call11.log
