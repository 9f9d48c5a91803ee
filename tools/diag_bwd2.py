"""Noise floor of gradient parity: the ORACLE network run with torch/cuDNN fp32 on the GPU vs on the CPU, next to
our TC path vs the CPU oracle.  Metrics: max-norm relative and L2 relative per weight tensor."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle.darknet_ref import RefDarknet
from oracle import region_loss_ref as RL
from singleshotpose_b200 import Darknet, RegionLoss, synth
from singleshotpose_b200.cfgs import write_cfg
cfg = write_cfg()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0); ref = RefDarknet(cfg); ref.train()
x, tgt = synth.images(B, seed=0), synth.targets(B, seed=1)
o = ref(x); l, _ = RL.region_loss_ref(o, tgt, 20); l.backward()
rg = {n: p.grad.clone() for n, p in ref.named_parameters()}
# oracle on the GPU through cuDNN
torch.manual_seed(0); refg = RefDarknet(cfg).cuda().train()
og = refg(x.cuda())
og.backward(torch.autograd.grad(RL.region_loss_ref(od := og.detach().cpu().requires_grad_(True), tgt, 20)[0], od)[0].cuda())
gg = {n: p.grad.detach().cpu() for n, p in refg.named_parameters()}
print("cudnn-vs-cpu logits rel %.2e" % float((og.detach().cpu() - o.detach()).abs().max() / o.detach().abs().max()))
torch.manual_seed(0); m = Darknet(cfg).cuda().train()
crit = RegionLoss(); crit.verbose = False
out = m(x.cuda()); crit(out, tgt, 20).backward()
g = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters()}
print("ours-vs-cpu logits rel %.2e" % float((out.detach().cpu() - o.detach()).abs().max() / o.detach().abs().max()))
mx = lambda a, b: float((a - b).abs().max() / b.abs().max())
l2 = lambda a, b: float((a - b).norm() / b.norm())
print("%-26s %9s %9s | %9s %9s" % ("param", "cudnn max", "cudnn l2", "ours max", "ours l2"))
for n in rg:
    if n.endswith("weight") and "bn" not in n:
        print("%-26s %9.2e %9.2e | %9.2e %9.2e" % (n, mx(gg[n], rg[n]), l2(gg[n], rg[n]), mx(g[n], rg[n]), l2(g[n], rg[n])))
