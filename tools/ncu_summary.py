"""Key metrics of every launch in an .ncu-rep (ncu --set full): python tools/ncu_summary.py rep > profiles/x.txt"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for n, r in enumerate(rows[2:]):
    print("== launch %d: %s" % (n, r[idx["Kernel Name"]]))
    for w in want:
        if w in idx:
            print("   %-70s %s %s" % (w, r[idx[w]], units[idx[w]]))
    # top warp-stall reasons (cycles a warp spends stalled per issued instruction, WarpStateStats)
    st = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try: st.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError: pass
    st.sort(reverse=True)
    if st:
        print("   stalls (warp-cycles per issued instruction): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:5]))
