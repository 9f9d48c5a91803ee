"""SASS opcode histogram per kernel of libssp_b200.so (cuobjdump -sass): the tcgen05 / TMA / TMEM evidence per kernel.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tile load (cp.async.bulk.tensor), LDTM = tcgen05.ld (TMEM -> registers),
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, RED/ATOM = global reductions, HMMA/FFMA/DFMA = legacy tensor / fp32 / fp64 pipes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "singleshotpose_b200", "csrc", "libssp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", ln)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        base = op.split(".")[0]
        if base in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "UTMAPF", "UTCATOMSWS"):
            hist[kern][op if base in ("UTCHMMA", "UTMALDG") and ".2CTA" in op else base] += 1
        elif base in ("SYNCS", "RED", "REDG", "ATOM", "ATOMG", "ATOMS", "HMMA", "FFMA", "DFMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT"):
            hist[kern][base] += 1
print("# %s" % os.path.relpath(lib, ROOT))
cols = ["_total", "UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D.2CTA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "RED", "REDG", "ATOM", "ATOMG", "HMMA", "FFMA", "DFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
seen = sorted({k for h in hist.values() for k in h})
cols = [c for c in cols if c in seen] + [c for c in seen if c not in cols]
wid = [max(7, len(c)) for c in cols]
print("%-46s " % "kernel" + " ".join("%*s" % (w, c) for w, c in zip(wid, cols)))
for k, h in hist.items():
    print("%-46s " % k.replace("void ", "").replace("ssp::", "")[:46] + " ".join("%*d" % (w, h.get(c, 0)) for w, c in zip(wid, cols)))
