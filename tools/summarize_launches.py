"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
seq = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1000.0)
    d = tot.setdefault(name, [0.0, 0]); d[0] += us; d[1] += 1
    seq.append((name, us))
total = sum(v[0] for v in tot.values())
print("# %d launches, %.3f ms total device time (cold-cache, serialised under ncu: compare SHARES)" % (len(seq), total / 1e3))
print("%-58s %8s %10s %7s %9s" % ("kernel", "launches", "total_us", "share", "avg_us"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%-58s %8d %10.1f %6.1f%% %9.1f" % (k[:58], v[1], v[0], 100 * v[0] / total, v[0] / v[1]))
if len(sys.argv) > 2:
    print("\n# sequence")
    for n, us in seq:
        print("%-58s %10.1f" % (n[:58], us))
