"""One warm-up + one profiled training step at batch B (cudaProfilerStart/Stop around the profiled step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from singleshotpose_b200 import Darknet, RegionLoss, FlatSGD, synth
from singleshotpose_b200.cfgs import write_cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = Darknet(write_cfg()).cuda().train()
crit = RegionLoss(); crit.verbose = False
opt = FlatSGD(m, lr=1e-6, momentum=0.9, weight_decay=0.03)
x, t = synth.images(B, seed=1).cuda(), synth.targets(B, seed=2).cuda()
def step():
    opt.zero_grad(); crit(m(x), t, 20).backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
