"""One GpuAugmenter batch (for ncu / compute-sanitizer): 8 LINEMOD-sized synthetic samples -> 416x416."""
import random
import sys

import torch

sys.path.insert(0, ".")
from singleshotpose_b200 import image as I, synth          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
samples = [synth.photo_sample(i) for i in range(n)]
imgs, masks, bgs = zip(*samples)
aug = I.GpuAugmenter("cuda", keep_u8=True)
x, params, u8 = aug(imgs, masks, bgs, (416, 416), rng=random.Random(0))
torch.cuda.synchronize()
print("ok", tuple(x.shape), float(x.mean()), int(u8.sum()))
