#!/bin/bash
# Round 2, call 13: cluster-of-4 grids sized by cudaOccupancyMaxActiveClusters: A/B again, launch list (grid sizes tell the cluster count).
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pnp 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kind']; print('$1  %.1f img/s  %.2f ms/step  e2e %.1f | fwd %.2f dgrad %.2f wgrad %.2f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['fwd']['ms_per_step'], r['dgrad']['ms_per_step'], r['wgrad']['ms_per_step']))"; }
{ run "default (clusters of 4)    ";
  SSP_TC2_CLUSTER=2 run "SSP_TC2_CLUSTER=2          ";
  run "default (repeat)           "; } | tee gpurun_out/r2_ab_call13.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "conv_gemm or wgrad_pair" 2>&1 | tail -2
SSP_OVERLAP=0 timeout 600 ncu --section SpeedOfLight --section LaunchStats --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum \
   --clock-control none --profile-from-start off -k regex:"conv_tc2|wgrad_tc2" -c 60 -o /tmp/r2_c4 python tools/one_step.py 64 > gpurun_out/r2_ncu13.log 2>&1
python tools/ncu_summary.py /tmp/r2_c4.ncu-rep | grep -E "^==|grid_size|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|xbar2l1tex" | paste - - - - - | sed 's/  */ /g' | cut -c1-300 > gpurun_out/r2_c4_summary.txt; head -70 gpurun_out/r2_c4_summary.txt
