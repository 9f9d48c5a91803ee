#!/bin/bash
# Round 2, call 16 (2 GPUs): data-parallel bench at N = 2 (bucketed all-reduce issued during backward), 2-rank NCCL tests; then sanitizers on GPU 0.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 2>&1 | tail -3
for b in 4 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --buckets $b 2>/dev/null | tail -1 > gpurun_out/r2_bench_n2_b$b.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_b$b.json')); print('N=2 buckets=$b  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done | tee gpurun_out/r2_n2_summary.txt
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_same_box.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_same_box.json')); print('N=1 (same box)  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a gpurun_out/r2_n2_summary.txt
CUDA_VISIBLE_DEVICES=0 bash tools/gpu_r2_sanitizers.sh
