"""Where does the tensor-core path lose precision at large K?  (a) fp16 subnormal lo parts, (b) accumulation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
from singleshotpose_b200 import _lib
from singleshotpose_b200._lib import call, ptr, stream_ptr
from test_gpu_kernels import flat_from_nchw, nchw_from_flat, _pack_w

for (cin, k) in [(128, 3), (512, 3), (1280, 3), (1280, 1)]:
    for wscale in (1.0, 64.0):
        N, H, W, cout = 1, 13, 13, 256
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, cin, H, W, generator=g)
        w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5 * wscale
        ref = F.conv2d(x.double(), w.double(), padding=(k - 1) // 2).float()
        xh, xl, rows = flat_from_nchw(x.cuda()); wh, wl, _ = _pack_w(w.cuda())
        res = {}
        for impl in (1, 0):
            y = torch.zeros(rows, cout, device="cuda")
            call("ssp_conv_gemm", impl, ptr(xh), ptr(xl), rows, cin, cin, ptr(wh), ptr(wl), cout, wh.shape[1], 0, 0,
                 N, H, W, k * k, cout, ptr(y), cout, rows, 0, None, None, None, stream_ptr())
            out = nchw_from_flat(y, N, cout, H, W).cpu()
            res[impl] = ((out - ref).abs().max() / ref.abs().max()).item(), ((out - ref).mean() / ref.abs().mean()).item()
        print("K=%5d wscale=%4g  simt relerr %.2e (bias %.1e)   tc relerr %.2e (bias %.1e)" % (cin * k * k, wscale, *res[1], *res[0]))
