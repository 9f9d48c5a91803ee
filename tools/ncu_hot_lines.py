"""Hottest SASS lines of one launch of an .ncu-rep captured with --set full --import-source on:
   python tools/ncu_hot_lines.py rep.ncu-rep <launch index> [top N]   -> address, samples, dominant stall reasons, instruction"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1] if rows and len(rows[0]) > 1 else "?")
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].replace(",", "").isdigit()]      # a second (CUDA-C) view repeats the header
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, " instructions", len(body))
order = sorted(range(len(body)), key=lambda k: -int(body[k][ix["# Samples"]] or 0))[:top]
for k in sorted(order):
    r = body[k]; n = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  %-28s exec %8s  %s" % (n, 100.0 * n / max(tot, 1), " ".join("%s:%d" % (w, c) for c, w in why if c), r[ix["Instructions Executed"]], r[ix["Source"]].strip()))
