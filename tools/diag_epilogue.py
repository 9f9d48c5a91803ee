"""Where does the time of the early conv layers go?  Times ssp_conv_gemm for epilogue modes F32 / STATS / none."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from singleshotpose_b200 import _lib
from singleshotpose_b200._lib import call, ptr, stream_ptr
B = 64
cases = [("L0 im2col 32->32 @416", 416, 32, 32, 1), ("L1 32->64 @208", 208, 32, 64, 9), ("L2 64->128 @104", 104, 64, 128, 9),
         ("L5 128->256 @52", 52, 128, 256, 9), ("L13 512->1024 @13", 13, 512, 1024, 9)]
for name, hw, cin, cout, taps in cases:
    rows = _lib.flat_alloc_rows(B, hw, hw)
    xh = torch.randn(rows, cin, device="cuda").half(); xl = (torch.randn(rows, cin, device="cuda") * 1e-3).half()
    kf = (taps * cin + 7) // 8 * 8
    wh = (torch.randn(cout, kf, device="cuda") * 0.05).half(); wl = (wh.float() * 1e-3).half()
    y = torch.zeros(rows, cout, device="cuda")
    ssum = torch.zeros(cout, dtype=torch.float64, device="cuda"); ssq = torch.zeros_like(ssum)
    res = []
    for impl in (0, 2):
        for epi in (0, 1, 3):
            if impl == 2 and epi == 3: continue
            def run():
                call("ssp_conv_gemm", impl, ptr(xh), ptr(xl), rows, cin, cin, ptr(wh), ptr(wl), cout, kf, 0, 0, B, hw, hw, taps, cout,
                     ptr(y), cout, rows, epi, None, ptr(ssum), ptr(ssq), stream_ptr())
            for _ in range(2): run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            res.append("impl%d/epi%d %.0fus" % (impl, epi, e0.elapsed_time(e1) * 200))
    print(name, " | ".join(res))
