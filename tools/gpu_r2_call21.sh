#!/bin/bash
# Round 2, call 21 (2 GPUs): e2e loop with the loss readback pipelined by one step -- N = 1 default bench line and N = 2 under torchrun.
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-cpu-baseline --no-pnp 2>/dev/null | tail -1 > gpurun_out/r2_bench21_n1.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench21_n1.json')); print('N=1  %.1f img/s  %.2f ms/step  e2e %.1f (%.2f ms/step)  loss %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('loss')))"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-cpu-baseline --no-pnp 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 (repeat)  %.1f img/s  %.2f ms/step  e2e %.1f (%.2f ms/step)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench21_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench21_n2.json')); print('N=2  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
