#!/bin/bash
# Round 2, call 10: ncu of the GEMM launches of one batch-64 step.  Reports stay in /tmp on the box (a --set full report of 66 launches is
# 110 MB, gpurun_out/ is capped at 64 MiB); only text summaries come back.
mkdir -p gpurun_out
K='regex:conv_tc2|wgrad_tc|conv_bandt|conv_tc_kernel'
SSP_OVERLAP=0 timeout 900 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables \
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_bytes.sum,smsp__cycles_active.avg,sm__cycles_elapsed.avg \
   --clock-control none --profile-from-start off -k "$K" -c 70 -o /tmp/r2_gemm_all python tools/one_step.py 64 > gpurun_out/r2_ncu10_all.log 2>&1
echo "all rc=$?"; python tools/ncu_summary.py /tmp/r2_gemm_all.ncu-rep > gpurun_out/r2_gemm_all_summary.txt
win() { # name skip count
  SSP_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" --launch-skip $2 --launch-count $3 -o /tmp/r2_w_$1 python tools/one_step.py 64 > gpurun_out/r2_ncu10_$1.log 2>&1
  python tools/ncu_summary.py /tmp/r2_w_$1.ncu-rep > gpurun_out/r2_gemm_$1_full.txt
  for i in $(seq 0 $(($3 - 1))); do echo "#### window $1 launch $i (global index $(($2 + i)))"; python tools/ncu_hot_lines.py /tmp/r2_w_$1.ncu-rep $i 16; done > gpurun_out/r2_gemm_$1_hot.txt 2>&1
}
win fwd_head 0 2        # block 2 (operand-swapped) and block 3 forward
win fwd_13 12 1         # first 13x13 3x3 forward
win bwd_tail 56 10      # blocks 6 ... 2: data gradient + weight gradient each
win bwd_13 24 4         # 13x13: dgrad / wgrad of the two widest layers
SSP_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"l0_gram|l0_fused|l0_bwd_kernel" -c 3 -o /tmp/r2_l0 python tools/one_step.py 64 > gpurun_out/r2_ncu10_l0.log 2>&1
python tools/ncu_summary.py /tmp/r2_l0.ncu-rep > gpurun_out/r2_l0_full.txt
for i in 0 1 2; do python tools/ncu_hot_lines.py /tmp/r2_l0.ncu-rep $i 14; done > gpurun_out/r2_l0_hot.txt 2>&1
ls -la gpurun_out | head -40; du -sh gpurun_out
