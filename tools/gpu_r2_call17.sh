#!/bin/bash
# Round 2, call 17: vectorised sgd_pack: tests + bench; collect-only count of the GPU suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -m gpu -q --timeout 300 -k "sgd or graph or load_weights" 2>&1 | tail -3
python -m pytest tests -m gpu --collect-only -q 2>/dev/null | tail -2
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "vectorised sgd_pack        "; run "vectorised sgd_pack (repeat)"; } | tee gpurun_out/r2_ab_call17.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sgd_pack -c 3 python tools/one_step.py 64 2>&1 | grep -E "sgd_pack|duration|bytes" | tail -8
