"""Oracle: ``get_region_boxes`` (reference utils.py:216-296) without the ``.cuda()`` calls.

Single-object decode: sigmoid on x0,y0,conf, grid offsets, softmax over the class
channel(s), then ONE box = the cell with the highest objectness over the WHOLE batch
(``max_conf`` is never reset per image; strict ``>`` keeps the first maximum in
(b, cy, cx) order).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import torch


def get_region_boxes_ref(output, num_classes, num_keypoints, only_objectness=1, validation=True):
    if output.dim() == 3:
        output = output.unsqueeze(0)
    batch, h, w = output.size(0), output.size(2), output.size(3)
    K = num_keypoints
    assert output.size(1) == 2 * K + 1 + num_classes
    out = output.view(batch, 2 * K + 1 + num_classes, h * w).transpose(0, 1).contiguous().view(
        2 * K + 1 + num_classes, batch * h * w)
    grid_x = torch.linspace(0, w - 1, w).repeat(h, 1).repeat(batch, 1, 1).view(batch * h * w)
    grid_y = torch.linspace(0, h - 1, h).repeat(w, 1).t().repeat(batch, 1, 1).view(batch * h * w)
    xs = [torch.sigmoid(out[0]) + grid_x] + [out[2 * j] + grid_x for j in range(1, K)]
    ys = [torch.sigmoid(out[1]) + grid_y] + [out[2 * j + 1] + grid_y for j in range(1, K)]
    det_confs = torch.sigmoid(out[2 * K])
    cls_confs = torch.softmax(out[2 * K + 1:2 * K + 1 + num_classes].transpose(0, 1), dim=1)
    cls_max_confs, cls_max_ids = torch.max(cls_confs, 1)
    max_conf = -float("inf")
    box = None
    for ind in range(batch * h * w):                     # (b, cy, cx) order, anchor_dim == 1
        conf = det_confs[ind] if only_objectness else det_confs[ind] * cls_max_confs[ind]
        if conf > max_conf:
            max_conf = conf
            box = []
            for j in range(K):
                box.append(xs[j][ind] / w)
                box.append(ys[j][ind] / h)
            box += [det_confs[ind], cls_max_confs[ind], cls_max_ids[ind]]
    return box
