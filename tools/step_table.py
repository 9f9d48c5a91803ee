"""Per-launch table from the raw-page CSV of a one-step ncu capture."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return d
print("%3s %-26s %6s %9s %8s %8s %7s %7s %7s %6s" % ("#", "kernel", "grid", "dur_us", "dramMB", "GB/s", "dram%", "l2%", "tens%", "occ%"))
tot = 0
for n, r in enumerate(rows[2:]):
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("ssp::", "")[:26]
    dur = g(r, "gpu__time_duration.sum"); u = units[ix["gpu__time_duration.sum"]]
    dur = dur / 1000 if u in ("ns", "nsecond") else (dur * 1000 if u.startswith("ms") else dur)
    dur0 = dur
    bps = g(r, "dram__bytes.sum.per_second"); ub = units[ix["dram__bytes.sum.per_second"]]
    bps *= {"byte/second": 1, "Kbyte/second": 1e3, "Mbyte/second": 1e6, "Gbyte/second": 1e9, "Tbyte/second": 1e12}.get(ub, 1)
    rd, wr = bps * dur0 * 1e-6 / 1e6, 0.0      # total DRAM MB during the launch
    tot += dur
    print("%3d %-26s %6d %9.1f %8.1f %8.1f %7.1f %7.1f %7.1f %6.1f" % (n, name, g(r, "launch__grid_size"), dur, rd, bps / 1e9,
          g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
          g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", -1) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in ix else g(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", -1),
          g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
print("total %.1f us" % tot)
