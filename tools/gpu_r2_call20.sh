#!/bin/bash
# Round 2, call 20 (2 GPUs): what the all-reduce of the 202 MB gradient buffer costs, and whether NCCL channel settings change it.
mkdir -p gpurun_out
cat > /tmp/ar.py <<'PY'
import os, torch, torch.distributed as dist
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
x = torch.ones(50547764, device="cuda")
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize()
for n in (50547764, 12000000, 3000000):
    y = x[:n]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_reduce(y)
    e1.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0: print("all_reduce %9d floats: %.3f ms  (%.0f GB/s algorithmic)" % (n, e0.elapsed_time(e1) / 10, n * 4 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9), flush=True)
dist.destroy_process_group()
PY
for env in "" "NCCL_MIN_NCHANNELS=32" "NCCL_MIN_NCHANNELS=32 NCCL_NTHREADS=512" "NCCL_ALGO=Ring NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=24"; do
  echo "== $env"; env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 /tmp/ar.py 2>&1 | grep all_reduce
done | tee gpurun_out/r2_nccl_allreduce_n2.txt
for env in "" "NCCL_MIN_NCHANNELS=32"; do
  env $env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 [$env]  %.1f img/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done | tee -a gpurun_out/r2_nccl_allreduce_n2.txt
