"""Condense tools/ncu_summary.py outputs into one line per launch: kernel, grid, duration, DRAM bytes, DRAM %, tensor %, warps active.

    python tools/step_from_full.py gemm.txt hbm.txt > profiles/r02_step_table.txt"""
import re
import sys

print("%-34s %9s %9s %9s %9s %7s %7s %7s %6s" % ("kernel", "grid", "dur_us", "dramR_MB", "dramW_MB", "dram%", "tens%el", "tens%ac", "warps%"))
for path in sys.argv[1:]:
    cur = None
    rows = []
    for ln in open(path):
        m = re.match(r"== launch (\d+): (.*)", ln)
        if m:
            cur = {"name": re.sub(r"\(.*", "", m.group(2)).replace("void ", "").replace("ssp::", "")}
            rows.append(cur)
            continue
        f = ln.split()
        if cur is None or len(f) < 2:
            continue
        try:
            v = float(f[1].replace(",", ""))
        except ValueError:
            continue
        unit = f[2] if len(f) > 2 else ""
        cur[f[0]] = (v, unit)

    def get(r, k, scale=1.0):
        if k not in r:
            return float("nan")
        v, u = r[k]
        mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        return v * mult * scale
    tot = 0.0
    for r in rows:
        d = get(r, "gpu__time_duration.sum")
        tot += d
        print("%-34s %9d %9.1f %9.1f %9.1f %7.1f %7.1f %7.1f %6.1f" % (
            r["name"][:34], int(get(r, "launch__grid_size")), d, get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum"),
            get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
    print("# %s: %d launches, %.1f us" % (path, len(rows), tot))
