"""Build libssp_b200.so in-tree with nvcc for sm_100a (no other architectures, no JIT cache)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["abi.cu", "conv_tc.cu", "conv_tc2.cu", "conv_band.cu", "conv_bandt.cu", "wgrad_tc.cu", "wgrad_tc2.cu", "conv_simt.cu", "conv0_direct.cu", "l0_fused.cu", "elementwise.cu", "sgd_pack.cu", "region.cu", "region_multi.cu", "pnp.cu", "augment.cu"]
# augment.cu restates Pillow's float/double pixel arithmetic bit for bit: no multiply-add contraction there
EXTRA = {"augment.cu": ["-fmad=false"]}
for _env, _macro in (("SSP_BN_MINBLOCKS", "SSP_BN_MINBLOCKS"), ("SSP_BN_UNITS", "BN_UNITS_PER_THREAD"),
                     ("SSP_BN_REDUCE_UNITS", "BWD_REDUCE_UNITS_PER_THREAD")):      # experiment knobs, see elementwise.cu
    if os.environ.get(_env):
        EXTRA.setdefault("elementwise.cu", []).append("-D%s=%d" % (_macro, int(os.environ[_env])))
LIB = os.path.join(HERE, "libssp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "..", "include", "ssp_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC, *FLAGS, *EXTRA.get(src, []), "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs])   # static cudart (nvcc default)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
