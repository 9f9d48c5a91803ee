import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.darknet_ref import RefDarknet
from oracle import region_loss_ref as RL
from singleshotpose_b200 import Darknet, RegionLoss, synth
from singleshotpose_b200.cfgs import write_cfg
cfg = write_cfg()
torch.manual_seed(0); ref = RefDarknet(cfg); ref.train()
x, tgt = synth.images(2, seed=0), synth.targets(2, seed=1)
o = ref(x); l, _ = RL.region_loss_ref(o, tgt, 20); l.backward()
rg = {n: p.grad.clone() for n, p in ref.named_parameters()}
def run(env):
    for k, v in env.items(): os.environ[k] = v
    torch.manual_seed(0); m = Darknet(cfg).cuda().train()
    crit = RegionLoss(); crit.verbose = False
    out = m(x.cuda()); crit(out, tgt, 20).backward()
    g = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters()}
    for k in env: os.environ.pop(k)
    return g
def rel(a, b): return float((a - b).abs().max() / b.abs().max())
g_tc = run({}); g_tc2 = run({})
g_simt = run({"SSP_CONV_IMPL": "simt", "SSP_WGRAD_IMPL": "simt"})
g_mix = run({"SSP_WGRAD_IMPL": "simt"})
print("%-28s %9s %9s %9s %9s %9s" % ("param", "tc/ref", "simt/ref", "tc/simt", "tc/tc2", "mix/simt"))
for n in rg:
    if "weight" in n and "bn" not in n or n.endswith("bn1.weight") or "conv23" in n:
        print("%-28s %9.2e %9.2e %9.2e %9.2e %9.2e" % (n, rel(g_tc[n], rg[n]), rel(g_simt[n], rg[n]), rel(g_tc[n], g_simt[n]), rel(g_tc[n], g_tc2[n]), rel(g_mix[n], g_simt[n])))
