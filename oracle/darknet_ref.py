"""Oracle: the yolo-pose network forward (and, through autograd, backward) in torch-CPU fp32.

Restates reference darknet.py: ``create_network`` (darknet.py:135-249) and ``forward``
(darknet.py:82-130) for the block types the two pose cfgs use (convolutional, maxpool 2/2,
route, reorg, region).  Modules are created in the reference's order with the reference's
constructor arguments so that a seeded default initialisation gives bit-identical
parameters (``torch.manual_seed(s); Darknet(cfg)`` == ``torch.manual_seed(s); RefDarknet(cfg)``).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _parse(cfgfile):
    # reference cfg.py:4-34
    blocks, cur = [], None
    for raw in open(cfgfile):
        line = raw.rstrip()
        if not line or line.startswith("#"):
            continue
        if line.startswith("["):
            if cur:
                blocks.append(cur)
            cur = {"type": line.strip("[]")}
            if cur["type"] == "convolutional":
                cur["batch_normalize"] = 0
        else:
            k, v = line.split("=")
            cur[k.strip()] = v.strip()
    blocks.append(cur)
    return blocks


def reorg_ref(x, stride=2):
    """reference darknet.py:16-35 (marvis ordering: out[b,(i*s+j)*C+c,h,w] = in[b,c,s*h+i,s*w+j])."""
    B, C, H, W = x.shape
    s = stride
    x = x.view(B, C, H // s, s, W // s, s).transpose(3, 4).contiguous()
    x = x.view(B, C, (H // s) * (W // s), s * s).transpose(2, 3).contiguous()
    x = x.view(B, C, s * s, H // s, W // s).transpose(1, 2).contiguous()
    return x.view(B, s * s * C, H // s, W // s)


class RefDarknet(nn.Module):
    def __init__(self, cfgfile):
        super().__init__()
        self.blocks = _parse(cfgfile)
        self.models = nn.ModuleList()
        prev = int(self.blocks[0].get("channels", 3))
        out_filters = []
        conv_id = 0
        for block in self.blocks[1:]:
            t = block["type"]
            if t == "convolutional":             # darknet.py:145-167
                conv_id += 1
                bn = int(block["batch_normalize"]); f = int(block["filters"]); k = int(block["size"])
                s = int(block["stride"]); pad = (k - 1) // 2 if int(block["pad"]) else 0
                m = nn.Sequential()
                if bn:
                    m.add_module("conv%d" % conv_id, nn.Conv2d(prev, f, k, s, pad, bias=False))
                    m.add_module("bn%d" % conv_id, nn.BatchNorm2d(f, eps=1e-4))
                else:
                    m.add_module("conv%d" % conv_id, nn.Conv2d(prev, f, k, s, pad))
                if block["activation"] == "leaky":
                    m.add_module("leaky%d" % conv_id, nn.LeakyReLU(0.1, inplace=True))
                prev = f
                out_filters.append(prev); self.models.append(m)
            elif t == "maxpool":                 # darknet.py:168-176
                self.models.append(nn.MaxPool2d(int(block["size"]), int(block["stride"])))
                out_filters.append(prev)
            elif t == "reorg":                   # darknet.py:198-202
                s = int(block["stride"]); prev = s * s * prev
                out_filters.append(prev); self.models.append(nn.Identity())
            elif t == "route":                   # darknet.py:203-213
                ind = len(self.models)
                layers = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
                prev = sum(out_filters[l] for l in layers)
                out_filters.append(prev); self.models.append(nn.Identity())
            elif t == "region":
                out_filters.append(prev); self.models.append(nn.Identity())
            else:
                raise ValueError("oracle does not model block type %r" % t)

    def forward(self, x):                        # darknet.py:82-130
        outputs = {}
        ind = -1
        for block in self.blocks[1:]:
            ind += 1
            t = block["type"]
            if t in ("convolutional", "maxpool"):
                x = self.models[ind](x)
            elif t == "reorg":
                x = reorg_ref(x, int(block["stride"]))
            elif t == "route":
                layers = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
                x = outputs[layers[0]] if len(layers) == 1 else torch.cat([outputs[l] for l in layers], 1)
            elif t == "region":
                continue
            outputs[ind] = x
        return x
