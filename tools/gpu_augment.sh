#!/bin/bash
# targeted GPU check of the image pipeline + the two new bench extras
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_augment.py -q --timeout 300 2>&1 | tail -15
timeout 300 python - <<'PY' 2>&1 | tail -12
import json, sys, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("augment", bench.augment_extra), ("multi", bench.multi_extra)):
    try:
        print(name, json.dumps(fn(dev, e0, e1)))
    except Exception as ex:
        import traceback; traceback.print_exc()
PY
