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
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for n, r in enumerate(rows[2:]):
    print("== launch %d: %s" % (n, r[idx["Kernel Name"]]))
    for w in want:
        if w in idx:
            print("   %-70s %s %s" % (w, r[idx[w]], units[idx[w]]))
