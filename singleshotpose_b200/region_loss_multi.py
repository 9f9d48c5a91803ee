"""``RegionLoss`` for the multi-object model -- drop-in for reference multi_obj_pose_estimation/region_loss_multi.py:94-189
(5 anchors, 13 classes: IoU anchor choice, masked corner MSE, confidence MSE, CrossEntropy(sum) on the class logits),
one kernel (ssp_region_loss_multi_fwd_bwd) including the gradient."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from ._lib import call, ptr, stream_ptr, SspError


class _RegionLossMultiFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, mod, epoch):
        nB, _, nH, nW = output.shape
        out = output.detach().contiguous().float()
        grad = torch.empty_like(out)
        acc = torch.empty(8, dtype=torch.float64, device=out.device)
        use_conf = 1 if epoch > mod.pretrain_num_epochs else 0
        step = len(mod.anchors) // mod.num_anchors
        anchors = (ctypes.c_float * len(mod.anchors))(*[float(a) for a in mod.anchors])
        call("ssp_region_loss_multi_fwd_bwd", ptr(out), ptr(target), ptr(grad), ptr(acc), nB, mod.num_keypoints, mod.num_classes,
             mod.num_anchors, nH, nW, ctypes.cast(anchors, ctypes.c_void_p), step, float(mod.coord_scale), float(mod.noobject_scale),
             float(mod.object_scale), float(mod.class_scale), float(mod.thresh), use_conf, 1.0, stream_ptr())
        ctx.save_for_backward(grad)
        mod._acc = acc
        loss = acc[0] + acc[1] + acc[6]
        if use_conf:
            loss = loss + acc[2]
        return loss.float()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


class RegionLoss(nn.Module):
    def __init__(self, num_keypoints=9, num_classes=13, anchors=[], num_anchors=5, pretrain_num_epochs=15):
        super().__init__()
        self.num_classes = num_classes
        self.anchors = anchors
        self.num_anchors = num_anchors
        self.anchor_step = len(anchors) / num_anchors
        self.num_keypoints = num_keypoints
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
        nl = 2 * self.num_keypoints + 3
        nch = (2 * self.num_keypoints + 1 + self.num_classes) * self.num_anchors
        if output.size(1) != nch:
            raise ValueError("output has %d channels, expected %d" % (output.size(1), nch))
        if len(self.anchors) < 2 * self.num_anchors:
            raise ValueError("anchors missing")
        if target.dim() != 2 or target.size(1) < 50 * nl or target.size(0) != output.size(0):
            raise ValueError("target must be (batch, 50*%d)" % nl)
        tgt = target.detach().to(device=output.device, dtype=torch.float32, non_blocking=True).contiguous()
        loss = _RegionLossMultiFn.apply(output, tgt, self, epoch)
        if self.verbose:
            a = self._acc.tolist()
            total = a[0] + a[1] + a[6] + (a[2] if epoch > self.pretrain_num_epochs else 0.0)
            print("%d: nGT %d, recall %d, proposals %d, loss: x %f, y %f, conf %f, cls %f, total %f" % (
                self.seen, int(a[3]), int(a[4]), int(a[5]), a[0], a[1], a[2], a[6], total))
        return loss

    def stats(self):
        a = self._acc.tolist()
        return dict(loss_x=a[0], loss_y=a[1], loss_conf=a[2], loss_cls=a[6], nGT=int(a[3]), nCorrect=int(a[4]), nProposals=int(a[5]))
