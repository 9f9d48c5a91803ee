"""One pass over every hand-written kernel at small shapes, for compute-sanitizer (racecheck / synccheck / memcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_kernels.py

Calls the parity tests' own bodies (tests/test_gpu_kernels.py), so every launch is also checked against its reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                            # noqa: E402
import test_gpu_kernels as T                            # noqa: E402
from singleshotpose_b200 import _lib                    # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
done = []


def run(name, fn, *a):
    fn(*a)
    torch.cuda.synchronize()
    done.append(name)


# CTA-pair kernel first (cross-CTA mbarriers, remote arrives, multicast commits), then the 1-CTA, band and wgrad kernels
for impl, tag in ((_lib.IMPL_TC2, "tc2"), (_lib.IMPL_TC, "tc"), (_lib.IMPL_BAND, "band")):
    if which in ("all", tag):
        for case in (T.CONV_CASES[2], T.CONV_CASES[4], T.CONV_CASES[1]):
            run("conv_%s%s" % (tag, case[:6]), T.test_conv_gemm_matches_torch, case, impl)
if which in ("all", "tc2"):
    run("dgrad_layout", T.test_conv_gemm_single_term_bf16_dgrad_layout)
if which in ("all", "wgrad"):
    for case in (T.WGRAD_CASES[1], T.WGRAD_CASES[2]):
        run("wgrad%s" % (case,), T.test_wgrad_gemm_matches_torch, case, _lib.IMPL_TC, _lib.FMT_F16)
    for case in (T.WGRAD_CASES[5], T.WGRAD_CASES[6]):       # merged-tap instructions (cin 32 / 128)
        run("wgrad_merged%s" % (case,), T.test_wgrad_gemm_matches_torch, case, _lib.IMPL_TC, _lib.FMT_F16)
if which in ("all", "wgrad2"):
    run("wgrad_pair", T.test_wgrad_pair_matches_single_cta_and_torch, (2, 13, 13, 256, 256, 3))
    run("wgrad_pair_c512", T.test_wgrad_pair_matches_single_cta_and_torch, (2, 26, 26, 256, 512, 3))
if which in ("all", "bandt"):                             # operand-swapped kernel: stacked hi/lo forward, single-term data gradient (fp32 + fp16 planes)
    for case in (T.BANDT_FWD[0], T.BANDT_FWD[2], T.BANDT_FWD[4]):
        run("bandt_fwd%s" % (case,), T.test_conv_bandt_forward_runs_and_matches_torch, case)
    for case in (T.BANDT_DGRAD[0], T.BANDT_DGRAD[2]):
        run("bandt_dgrad%s" % (case,), T.test_conv_bandt_dgrad_runs_and_matches_torch, case, True)
if which in ("all", "l0"):                                # blocks 0-1 as one unit
    run("l0_fused", T.test_l0_fused_blocks_match_torch, (2, 34, 70))
if which in ("all", "tc2"):
    run("tc2_f16_plane", T.test_conv_tc2_dgrad_fp16_plane, (2, 13, 13, 256, 512, 3))
if which in ("all", "misc"):
    run("bn_pool", T.test_bn_apply_and_backward, _lib.ROUTE_POOL, 32, True)
    run("bn_direct", T.test_bn_apply_and_backward, _lib.ROUTE_DIRECT, 256, False)
    run("conv0", T.test_conv0_direct_matches_torch, (2, 32, 40))
    run("sgd_pack", T.test_sgd_pack_step_matches_separate_kernels)
print("sanitize_kernels: ran %d kernel checks: %s" % (len(done), ", ".join(done)))
