#!/bin/bash
# Round 2, call 29 (2 GPUs): does leaving SMs free for NCCL (SSP_SM_LIMIT) let the gradient buckets overlap with the backward GEMMs?
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>gpurun_out/r2_c29.err | tail -1 > gpurun_out/r2_c29.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_c29.json')); print('$label  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" || tail -5 gpurun_out/r2_c29.err
}
run "N=2 default                      " A=1
run "N=2 SM_LIMIT=144 NCCL 4 channels " SSP_SM_LIMIT=144 NCCL_MAX_NCHANNELS=4 NCCL_MIN_NCHANNELS=4
run "N=2 SM_LIMIT=140 NCCL 8 channels " SSP_SM_LIMIT=140 NCCL_MAX_NCHANNELS=8 NCCL_MIN_NCHANNELS=8
run "N=2 SM_LIMIT=146 NCCL 2 channels " SSP_SM_LIMIT=146 NCCL_MAX_NCHANNELS=2 NCCL_MIN_NCHANNELS=2
run "N=2 SM_LIMIT=148 NCCL 4 channels " NCCL_MAX_NCHANNELS=4 NCCL_MIN_NCHANNELS=4
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 (same box)  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
CUDA_VISIBLE_DEVICES=0 SSP_SM_LIMIT=144 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 SM_LIMIT=144  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
