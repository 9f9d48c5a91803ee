import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from singleshotpose_b200 import Darknet, RegionLoss, FlatSGD, synth, GraphedTrainStep
from singleshotpose_b200.cfgs import write_cfg
cfg = write_cfg()
torch.manual_seed(4)
a = Darknet(cfg).cuda().train(); b = copy.deepcopy(a)
x, tgt = synth.images(2, seed=8), synth.targets(2, seed=9)
crit = RegionLoss(); crit.verbose = False
opt_a = FlatSGD(a, lr=1e-5, momentum=0.9, weight_decay=0.01); opt_b = FlatSGD(b, lr=1e-5, momentum=0.9, weight_decay=0.01)
cb = RegionLoss(); cb.verbose = False
g = GraphedTrainStep(b, cb, opt_b, (2, 3, 416, 416), (2, 1050), 20, torch.device("cuda"), warmup=1)
state = copy.deepcopy(b.state_dict())
g.x.copy_(x); g.t.copy_(tgt); g.capture()
b.load_state_dict(state); opt_b._v.zero_()
print("param diff after restore", max(float((p - q).abs().max()) for p, q in zip(a.parameters(), b.parameters())))
for it in range(3):
    opt_a.zero_grad(); la = crit(a(x.cuda()), tgt, 20); la.backward(); opt_a.step()
    lb = g(x.pin_memory(), tgt.pin_memory())
    torch.cuda.synchronize()
    pd = max(float((p - q).abs().max() / q.abs().max()) for p, q in zip(a.parameters(), b.parameters()))
    gd = float((a._engine.flat_grads - b._engine.flat_grads).norm() / a._engine.flat_grads.norm())
    print(it, float(la), float(lb), "param rel diff", pd, "grad rel diff", gd)
