"""``Darknet`` -- drop-in for reference darknet.py:59-394 whose forward/backward run on libssp_b200.so.

Same constructor, attributes, parameter names (``models.<i>.conv<j>.weight``, ``bn<j>.weight/bias``), seeded
initialisation and ``.weights`` file format as the reference, so ``train.py`` / ``valid.py`` work unchanged
(see singleshotpose_b200/dropin/).  The nn.Conv2d / nn.BatchNorm2d modules only HOLD the parameters; they are
never called: ``forward`` hands the whole stack to the Engine (engine.py) behind one autograd Function.
CPU tensors are rejected -- there is no fallback path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .cfg import parse_cfg, print_cfg, load_conv, load_conv_bn, save_conv, save_conv_bn
from .engine import Engine


class Reorg(nn.Module):
    """Placeholder for reference darknet.py:16-35; the data movement is fused into the producer's BN-apply kernel."""

    def __init__(self, stride=2):
        super().__init__()
        self.stride = stride


class EmptyModule(nn.Module):
    """route blocks (darknet.py:51-56)."""


class _MaxPoolMarker(nn.Module):
    def __init__(self, size, stride):
        super().__init__()
        self.kernel_size, self.stride = size, stride


class _DarknetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, needs_grad, x, *params):
        eng = model._engine
        out, bufs, gen = eng.forward(x, train_bn=model.training, keep_for_backward=needs_grad)
        ctx.eng, ctx.bufs, ctx.gen, ctx.nparams = eng, bufs, gen, len(params)
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.eng
        aliased = [p for p in ctx.params if p.grad is not None and p.grad.data_ptr() == eng.grad_view(p).data_ptr()]
        prev = eng.flat_grads.clone() if aliased else None
        eng.backward(ctx.bufs, ctx.gen, grad_out)
        if prev is not None:
            # p.grad already IS the flat gradient buffer (zero_grad(set_to_none=False) / accumulation): add in place
            eng.flat_grads.add_(prev)
            return (None, None, None) + tuple(None for _ in range(ctx.nparams))
        return (None, None, None) + tuple(eng.grad_view(p) for p in ctx.params)


class Darknet(nn.Module):
    _region_loss_cls = None        # darknet_multi.Darknet overrides this with the multi-object head

    def __init__(self, cfgfile):
        super().__init__()
        self.blocks = parse_cfg(cfgfile)
        self.models = self.create_network(self.blocks)
        self.loss = self.models[len(self.models) - 1]
        net = self.blocks[0]
        self.width = int(net["width"])
        self.height = int(net["height"])
        self.test_width = int(net.get("test_width", net["width"]))
        self.test_height = int(net.get("test_height", net["height"]))
        self.num_keypoints = int(net.get("num_keypoints", 9))
        self.loss.num_keypoints = self.num_keypoints if hasattr(self.loss, "num_keypoints") else None
        if self.blocks[-1]["type"] == "region":
            self.anchors = self.loss.anchors
            self.num_anchors = self.loss.num_anchors
            self.anchor_step = self.loss.anchor_step
            self.num_classes = self.loss.num_classes
        self.header = torch.IntTensor([0, 0, 0, 0])
        self.seen = 0
        self.iter = 0
        self._engine = Engine(self)

    # ---- construction: same module order / ctor args as darknet.py:135-249 so that seeded init is identical ----
    def create_network(self, blocks):
        from .region_loss import RegionLoss as _SingleLoss
        RegionLoss = self._region_loss_cls or _SingleLoss
        models = nn.ModuleList()
        prev_filters = 3
        out_filters = []
        conv_id = 0
        for block in blocks:
            t = block["type"]
            if t == "net":
                prev_filters = int(block.get("channels", 3))
                continue
            if t == "convolutional":
                conv_id += 1
                bn = int(block["batch_normalize"])
                filters, k, stride = int(block["filters"]), int(block["size"]), int(block["stride"])
                pad = (k - 1) // 2 if int(block["pad"]) else 0
                model = nn.Sequential()
                if bn:
                    model.add_module("conv%d" % conv_id, nn.Conv2d(prev_filters, filters, k, stride, pad, bias=False))
                    model.add_module("bn%d" % conv_id, nn.BatchNorm2d(filters, eps=1e-4))
                else:
                    model.add_module("conv%d" % conv_id, nn.Conv2d(prev_filters, filters, k, stride, pad))
                if block["activation"] == "leaky":
                    model.add_module("leaky%d" % conv_id, nn.LeakyReLU(0.1, inplace=True))
                prev_filters = filters
                out_filters.append(prev_filters)
                models.append(model)
            elif t == "maxpool":
                out_filters.append(prev_filters)
                models.append(_MaxPoolMarker(int(block["size"]), int(block["stride"])))
            elif t == "reorg":
                stride = int(block["stride"])
                prev_filters = stride * stride * prev_filters
                out_filters.append(prev_filters)
                models.append(Reorg(stride))
            elif t == "route":
                ind = len(models)
                layers = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
                prev_filters = sum(out_filters[l] for l in layers)
                out_filters.append(prev_filters)
                models.append(EmptyModule())
            elif t == "region":
                loss = RegionLoss()
                anchors = block["anchors"].split(",")
                loss.anchors = [] if anchors == [""] else [float(i) for i in anchors]
                loss.num_classes = int(block["classes"])
                loss.num_anchors = int(block["num"])
                loss.anchor_step = len(loss.anchors) // loss.num_anchors
                loss.object_scale = float(block["object_scale"])
                loss.noobject_scale = float(block["noobject_scale"])
                loss.class_scale = float(block["class_scale"])
                loss.coord_scale = float(block["coord_scale"])
                out_filters.append(prev_filters)
                models.append(loss)
            else:
                raise NotImplementedError("block type %r is outside the hot path" % t)
        return models

    def forward(self, x):
        """(B,3,H,W) CUDA fp32 -> (B, (2K+1+C)*A, H/32, W/32) raw output of the last conv; the region block is skipped
        exactly as in darknet.py:119-120."""
        self.loss = None
        params = tuple(self.parameters())
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)   # grad mode is off inside Function.forward
        return _DarknetFn.apply(self, needs_grad, x, *params)

    def print_network(self):
        print_cfg(self.blocks)

    # ---- Darknet .weights I/O: int32[4] header + fp32 stream (darknet.py:251-394, cfg.py:153-190) ----
    def _read(self, weightfile):
        with open(weightfile, "rb") as fp:
            header = np.fromfile(fp, count=4, dtype=np.int32)
            buf = np.fromfile(fp, dtype=np.float32)
        self.header = torch.from_numpy(header)
        self.seen = self.header[3]
        return buf

    def _load(self, buf, nblocks):
        start, ind = 0, -2
        for block in self.blocks[:nblocks]:
            if start >= buf.size:
                break
            ind += 1
            if block["type"] == "convolutional":
                model = self.models[ind]
                if int(block["batch_normalize"]):
                    start = load_conv_bn(buf, start, model[0], model[1])
                else:
                    start = load_conv(buf, start, model[0])
        self._engine.invalidate_packed_weights()      # the fp16 operand copies of the conv weights are stale now
        return start

    def load_weights(self, weightfile):
        self._load(self._read(weightfile), len(self.blocks))

    def load_weights_until_last(self, weightfile):
        """all blocks except the last two (last conv + region), darknet.py:299-347"""
        self._load(self._read(weightfile), len(self.blocks) - 2)

    def save_weights(self, outfile, cutoff=0):
        if cutoff <= 0:
            cutoff = len(self.blocks) - 1
        with open(outfile, "wb") as fp:
            self.header[3] = int(self.seen)
            self.header.numpy().tofile(fp)
            ind = -1
            for block_id in range(1, cutoff + 1):
                ind += 1
                block = self.blocks[block_id]
                if block["type"] == "convolutional":
                    model = self.models[ind]
                    if int(block["batch_normalize"]):
                        save_conv_bn(fp, model[0], model[1])
                    else:
                        save_conv(fp, model[0])
