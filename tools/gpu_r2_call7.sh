#!/bin/bash
# Round 2, call 7: the whole GPU suite on the new defaults, smoke(), A/B of the split-K data gradient, default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_t_all7.log 2>&1; tail -8 gpurun_out/r2_t_all7.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke7.log 2>&1; tail -3 gpurun_out/r2_smoke7.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "default                    ";
  SSP_DGRAD_SPLITK=1 run "SSP_DGRAD_SPLITK=1         ";
  SSP_OVERLAP=0 run "SSP_OVERLAP=0 (wgrad inline)";
  run "default (repeat)           "; } | tee gpurun_out/r2_ab_call7.log
timeout 900 python bench.py > gpurun_out/r2_bench_default7.json 2> gpurun_out/r2_bench_default7.err; tail -c 1500 gpurun_out/r2_bench_default7.json
