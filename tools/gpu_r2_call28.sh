#!/bin/bash
# Round 2, call 28 (4 GPUs): the data-parallel bench at N = 4 (bucketed all-reduce inside the captured step), and N = 1 on the same box.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/r2_bench_n4.err | tail -1 > gpurun_out/r2_bench_n4.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n4.json')); print('N=4  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/r2_bench_n4.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 (same box)  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
