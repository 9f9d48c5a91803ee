#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 600 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
