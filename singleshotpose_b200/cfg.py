"""Network-description parsing and Darknet ``.weights`` (de)serialisation.

Host-side mirror of reference cfg.py: ``parse_cfg`` (cfg.py:4-34), ``print_cfg``
(cfg.py:36-151), ``load_conv/_bn/_fc`` and ``save_conv/_bn/_fc`` (cfg.py:153-201).
Same names, arguments and return values; re-implemented (no reference code is imported).
"""
from __future__ import annotations

import numpy as np
import torch


def parse_cfg(cfgfile):
    """cfg text -> list of dict blocks.  Semantics of reference cfg.py:4-34: sections start
    with ``[name]``; ``#`` and blank lines skipped; a key literally named ``type`` is stored
    as ``_type``; ``[convolutional]`` blocks default ``batch_normalize`` to 0; values stay
    strings."""
    blocks = []
    cur = None
    with open(cfgfile, "r") as fp:
        for raw in fp:
            line = raw.rstrip()
            if not line or line[0] == "#":
                continue
            if line[0] == "[":
                if cur:
                    blocks.append(cur)
                cur = {"type": line.lstrip("[").rstrip("]")}
                if cur["type"] == "convolutional":
                    cur["batch_normalize"] = 0
                continue
            key, value = line.split("=")
            key = key.strip()
            cur["_type" if key == "type" else key] = value.strip()
    if cur:
        blocks.append(cur)
    return blocks


def layer_shapes(blocks, width=None, height=None):
    """Walk *blocks* and return, per non-net block, (kind, in_ch, out_ch, in_h, in_w, out_h, out_w, extra).
    Shared by print_cfg and the execution-plan builder."""
    net = blocks[0]
    w = int(net["width"]) if width is None else width
    h = int(net["height"]) if height is None else height
    c = int(net.get("channels", 3))
    outs = []
    ind = -1
    for block in blocks[1:]:
        ind += 1
        t = block["type"]
        if t == "convolutional":
            k = int(block["size"]); s = int(block["stride"])
            pad = (k - 1) // 2 if int(block["pad"]) else 0
            f = int(block["filters"])
            ow = (w + 2 * pad - k) // s + 1
            oh = (h + 2 * pad - k) // s + 1
            outs.append(("conv", c, f, h, w, oh, ow, dict(size=k, stride=s, pad=pad)))
            c, h, w = f, oh, ow
        elif t == "maxpool":
            s = int(block["stride"]); k = int(block["size"])
            outs.append(("max", c, c, h, w, h // s, w // s, dict(size=k, stride=s)))
            h, w = h // s, w // s
        elif t == "reorg":
            s = int(block["stride"])
            outs.append(("reorg", c, c * s * s, h, w, h // s, w // s, dict(stride=s)))
            c, h, w = c * s * s, h // s, w // s
        elif t == "route":
            layers = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
            if len(layers) == 1:
                _, _, oc, _, _, oh, ow, _ = outs[layers[0]]
            else:
                oc = outs[layers[0]][2] + outs[layers[1]][2]
                oh, ow = outs[layers[0]][5], outs[layers[0]][6]
                assert (oh, ow) == (outs[layers[1]][5], outs[layers[1]][6])
            outs.append(("route", c, oc, h, w, oh, ow, dict(layers=layers)))
            c, h, w = oc, oh, ow
        elif t == "region":
            outs.append(("detection", c, c, h, w, h, w, {}))
        else:
            outs.append((t, c, c, h, w, h, w, {}))
    return outs


def print_cfg(blocks):
    """Layer table in the reference's column format (cfg.py:36-151)."""
    print("layer     filters    size              input                output")
    for ind, (kind, ic, oc, ih, iw, oh, ow, ex) in enumerate(layer_shapes(blocks)):
        if kind == "conv":
            print("%5d %-6s %4d  %d x %d / %d   %3d x %3d x%4d   ->   %3d x %3d x%4d" % (
                ind, "conv", oc, ex["size"], ex["size"], ex["stride"], iw, ih, ic, ow, oh, oc))
        elif kind == "max":
            print("%5d %-6s       %d x %d / %d   %3d x %3d x%4d   ->   %3d x %3d x%4d" % (
                ind, "max", ex["size"], ex["size"], ex["stride"], iw, ih, ic, ow, oh, oc))
        elif kind == "reorg":
            print("%5d %-6s             / %d   %3d x %3d x%4d   ->   %3d x %3d x%4d" % (
                ind, "reorg", ex["stride"], iw, ih, ic, ow, oh, oc))
        elif kind == "route":
            print("%5d %-6s %s" % (ind, "route", " ".join(str(l) for l in ex["layers"])))
        elif kind == "detection":
            print("%5d %-6s" % (ind, "detection"))
        else:
            print("unknown type %s" % kind)


def _take(buf, start, tensor):
    n = tensor.numel()
    with torch.no_grad():          # a tracked in-place write: bumps tensor._version, which the engine's operand-plane cache keys on
        tensor.copy_(torch.from_numpy(np.ascontiguousarray(buf[start:start + n])).view(tensor.shape))
    return start + n


def load_conv(buf, start, conv_model):
    """bias then weight (reference cfg.py:153-158)."""
    start = _take(buf, start, conv_model.bias)
    return _take(buf, start, conv_model.weight)


def load_conv_bn(buf, start, conv_model, bn_model):
    """bn.bias, bn.weight, running_mean, running_var, conv.weight (reference cfg.py:169-176)."""
    for t in (bn_model.bias, bn_model.weight, bn_model.running_mean, bn_model.running_var):
        start = _take(buf, start, t)
    return _take(buf, start, conv_model.weight)


def load_fc(buf, start, fc_model):
    start = _take(buf, start, fc_model.bias)
    return _take(buf, start, fc_model.weight)


def _dump(fp, tensor):
    tensor.detach().to("cpu", torch.float32).contiguous().numpy().tofile(fp)


def save_conv(fp, conv_model):
    _dump(fp, conv_model.bias); _dump(fp, conv_model.weight)


def save_conv_bn(fp, conv_model, bn_model):
    for t in (bn_model.bias, bn_model.weight, bn_model.running_mean, bn_model.running_var, conv_model.weight):
        _dump(fp, t)


def save_fc(fp, fc_model):
    _dump(fp, fc_model.bias); _dump(fp, fc_model.weight)
