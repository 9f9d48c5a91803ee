#!/bin/bash
# Round 2, call 4: fused blocks 0-1 (csrc/l0_fused.cu): kernel test, network tests, same-box A/B against the direct path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -k "l0_fused" > gpurun_out/r2_t_call4a.log 2>&1; tail -25 gpurun_out/r2_t_call4a.log
timeout 1200 python -m pytest tests/test_gpu_network.py tests/test_gpu_multi.py -m gpu -q --timeout 900 > gpurun_out/r2_t_call4b.log 2>&1; tail -25 gpurun_out/r2_t_call4b.log
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
run "L0 fused (default)      " | tee gpurun_out/r2_ab_call4.log
SSP_L0=direct run "L0 direct (round-1 path)" | tee -a gpurun_out/r2_ab_call4.log
run "L0 fused (repeat)       " | tee -a gpurun_out/r2_ab_call4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 500 --csv --log-file gpurun_out/r2_launches4.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pnp --no-graph > gpurun_out/r2_bench_under_ncu4.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches4.csv seq > gpurun_out/r2_launches4.txt 2>&1; head -34 gpurun_out/r2_launches4.txt
