"""Same-box A/B of a library build: per-layer GEMM timings (forward 3-term, dgrad 1-term, wgrad)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from singleshotpose_b200 import _lib
from singleshotpose_b200._lib import call, ptr, stream_ptr
B = 64
def timeit(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n
cases = [("L1 32->64 @208", 208, 32, 64, 9), ("L2 64->128 @104", 104, 64, 128, 9), ("L5 128->256 @52", 52, 128, 256, 9),
         ("L8 256->512 @26", 26, 256, 512, 9), ("L13 512->1024 @13", 13, 512, 1024, 9), ("L14 1x1 1024->512 @13", 13, 1024, 512, 1)]
for name, hw, cin, cout, taps in cases:
    rows = _lib.flat_alloc_rows(B, hw, hw)
    xh = torch.randn(rows, cin, device="cuda").half(); xl = (torch.randn(rows, cin, device="cuda") * 1e-3).half()
    kf = (taps * cin + 7) // 8 * 8
    wh = (torch.randn(cout, kf, device="cuda") * 0.05).half(); wl = (wh.float() * 1e-3).half()
    y = torch.zeros(rows, cout, device="cuda")
    ssum = torch.zeros(cout, dtype=torch.float64, device="cuda"); ssq = torch.zeros_like(ssum)
    dyp = torch.randn(rows, cout, device="cuda").half()
    wd = (torch.randn(cin, (taps * cout + 7) // 8 * 8, device="cuda") * 0.05).half()
    dx = torch.zeros(rows, cin, device="cuda")
    dw = torch.zeros(cout, taps, cin, device="cuda")
    out = []
    for impl in ([0, 2, 3] if taps == 9 else [0, 2]):
        t = timeit(lambda: call("ssp_conv_gemm", impl, ptr(xh), ptr(xl), rows, cin, cin, ptr(wh), ptr(wl), cout, kf, 0, 0, B, hw, hw, taps, cout,
                                ptr(y), cout, rows, 1, None, ptr(ssum), ptr(ssq), stream_ptr()))
        out.append("fwd impl%d %.0f" % (impl, t))
    for impl in ([0, 2, 3] if taps == 9 else [0, 2]):
        t = timeit(lambda: call("ssp_conv_gemm", impl, ptr(dyp), None, rows, cout, cout, ptr(wd), None, cin, wd.shape[1], 0, 0, B, hw, hw, taps, cin,
                                ptr(dx), cin, rows, 0, None, None, None, stream_ptr()))
        out.append("dgrad impl%d %.0f" % (impl, t))
    t = timeit(lambda: call("ssp_wgrad_gemm", 0, ptr(dyp), rows, cout, cout, 0, ptr(xh), rows, cin, cin, 0, B, hw, hw, taps, ptr(dw), cin, cin, 1.0, stream_ptr()))
    out.append("wgrad %.0f" % t)
    print(name, "|", " ".join(out))
