"""writes the golden sigma = 0 / 5 px problems as the binary input of tools/probes/pnp_probe.cu"""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1]
a = np.load(os.path.join(root, "tests", "golden", "pnp.npz")); b = np.load(os.path.join(root, "tests", "golden", "pnp_noise.npz"))
uv = np.concatenate([a["uv_s0"], b["uv_s5"]]).astype(np.float32)
with open(out, "wb") as f:
    np.array([uv.shape[0], 9], np.int32).tofile(f); a["K"].astype(np.float32).tofile(f); a["P3"].astype(np.float32).tofile(f); uv.tofile(f)
