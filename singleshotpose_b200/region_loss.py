"""``RegionLoss`` -- drop-in for reference region_loss.py:80-175 (single-object head).

One kernel (ssp_region_loss_fwd_bwd) does activation, corner decode, build_targets, the masked MSE terms, the
counters and the gradient w.r.t. the raw network output; ``target`` may arrive on the CPU as in train.py:82-97.
The reference's per-iteration log line (region_loss.py:173) is kept (set ``verbose=False`` to skip the
device->host read it needs).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import call, ptr, stream_ptr, SspError


class _RegionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, mod, epoch):
        nB, _, nH, nW = output.shape
        out = output.detach().contiguous().float()
        grad = torch.empty_like(out)
        acc = torch.empty(8, dtype=torch.float64, device=out.device)
        use_conf = 1 if epoch > mod.pretrain_num_epochs else 0
        call("ssp_region_loss_fwd_bwd", ptr(out), ptr(target), ptr(grad), ptr(acc), nB, mod.num_keypoints, mod.num_classes,
             nH, nW, float(mod.coord_scale), float(mod.noobject_scale), float(mod.object_scale), float(mod.thresh),
             use_conf, 1.0, stream_ptr())
        ctx.save_for_backward(grad)
        mod._acc = acc
        loss = acc[0] + acc[1]
        if use_conf:
            loss = loss + acc[2]
        return loss.float()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


class RegionLoss(nn.Module):
    def __init__(self, num_keypoints=9, num_classes=1, anchors=[], num_anchors=1, pretrain_num_epochs=15):
        super().__init__()
        self.num_classes = num_classes
        self.num_anchors = num_anchors
        self.num_keypoints = num_keypoints
        self.anchors = anchors
        self.anchor_step = len(anchors) // num_anchors if num_anchors else 0
        self.coord_scale = 1
        self.noobject_scale = 1
        self.object_scale = 5
        self.class_scale = 1
        self.thresh = 0.6
        self.seen = 0
        self.pretrain_num_epochs = pretrain_num_epochs
        self.verbose = True
        self._acc = None

    def forward(self, output, target, epoch):
        if not output.is_cuda:
            raise SspError("RegionLoss runs on CUDA tensors only (no CPU fallback)")
        if self.num_anchors != 1:
            raise NotImplementedError("multi-anchor RegionLoss (region_loss_multi.py) is not built yet")
        nl = 2 * self.num_keypoints + 3
        if target.dim() != 2 or target.size(1) != 50 * nl or target.size(0) != output.size(0):
            # the kernel strides rows by 50*nl and reads the first ground truth (the reference supports exactly one per image:
            # a second one is a broadcast error at region_loss.py:39, none an IndexError at :40 -- SURVEY 8a/a10)
            raise ValueError("target must be (batch, 50*%d), got %s" % (nl, tuple(target.shape)))
        tgt = target.detach().to(device=output.device, dtype=torch.float32, non_blocking=True).contiguous()
        loss = _RegionLossFn.apply(output, tgt, self, epoch)
        if self.verbose:
            a = self._acc.tolist()
            total = a[0] + a[1] + (a[2] if epoch > self.pretrain_num_epochs else 0.0)
            print("%d: nGT %d, recall %d, proposals %d, loss: x %f, y %f, conf %f, total %f" % (
                self.seen, int(a[3]), int(a[4]), int(a[5]), a[0], a[1], a[2], total))
        return loss

    def stats(self):
        """dict of the last call's loss parts and counters (one device->host read)."""
        a = self._acc.tolist()
        return dict(loss_x=a[0], loss_y=a[1], loss_conf=a[2], nGT=int(a[3]), nCorrect=int(a[4]), nProposals=int(a[5]))
