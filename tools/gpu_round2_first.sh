#!/bin/bash
# First GPU call of round 2 (≈2 min): answers the three open questions of profiles/r01_experiments.md in one go.
#  1. does the experimental CTA-pair weight gradient (csrc/wgrad_tc2.cu) pass its parity test?
#  2. same-box A/B of the training step with SSP_WGRAD_IMPL=tc vs tc2 (graph replay, no CPU baseline, no extras)
#     (+ one run each with the experimental tiled weight re-pack, SSP_PACK=v2, and the BN-backward/dgrad fusion, SSP_FUSE_BNBWD=1)
#  3. TMA fill rate per SM with own tiles / shared tiles / cluster multicast (tools/probes/mc_probe.cu)
mkdir -p gpurun_out
SSP_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -q -k "wgrad_pair or pack_weights_v2 or fused_bn_backward or other_resolutions_match" --timeout 300 2>&1 | tail -6
for impl in tc tc2 tc tc2 v2pack bnfuse sgdplain; do
  w=$impl
  if [ $impl = v2pack ]; then export SSP_PACK=v2; w=tc; fi
  if [ $impl = bnfuse ]; then unset SSP_PACK; export SSP_FUSE_BNBWD=1; w=tc; fi
  if [ $impl = sgdplain ]; then unset SSP_FUSE_BNBWD; export SSP_SGD_FUSED=0; w=tc; fi
  SSP_WGRAD_IMPL=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null > gpurun_out/ab_$impl.json
  python - "$impl" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_%s.json" % sys.argv[1]))
w = d["roofline"]["per_kind"]["wgrad"]
print("wgrad=%s  %.1f img/s  %.2f ms/step  wgrad %.2f ms/step (%.0f TFLOP/s)" % (sys.argv[1], d["value"], d["ms_per_step"], w["ms_per_step"], w["tflops"]))
PY
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I singleshotpose_b200/csrc -o /tmp/mc_probe tools/probes/mc_probe.cu 2>/dev/null && timeout 60 /tmp/mc_probe
