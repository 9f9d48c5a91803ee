"""Oracle: ``get_multi_region_boxes`` (reference multi_obj_pose_estimation/utils_multi.py:266-382) without ``.cuda()``.
Keeps every (cell, anchor) whose confidence exceeds ``conf_thresh`` in (cy, cx, anchor) order, plus a fallback box for
``correspondingclass`` (running maxima: ``max_conf`` reset per image, ``max_cls_conf`` never reset).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch


def get_multi_region_boxes_ref(output, conf_thresh, num_classes, num_keypoints, anchors, num_anchors, correspondingclass,
                               only_objectness=1, validation=False):
    if output.dim() == 3:
        output = output.unsqueeze(0)
    batch, h, w = output.size(0), output.size(2), output.size(3)
    K, nC, nA = num_keypoints, num_classes, num_anchors
    assert output.size(1) == (2 * K + 1 + nC) * nA
    out = output.view(batch * nA, 2 * K + 1 + nC, h * w).transpose(0, 1).contiguous().view(2 * K + 1 + nC, batch * nA * h * w)
    grid_x = torch.linspace(0, w - 1, w).repeat(h, 1).repeat(batch * nA, 1, 1).view(batch * nA * h * w)
    grid_y = torch.linspace(0, h - 1, h).repeat(w, 1).t().repeat(batch * nA, 1, 1).view(batch * nA * h * w)
    xs = [torch.sigmoid(out[0]) + grid_x] + [out[2 * j] + grid_x for j in range(1, K)]
    ys = [torch.sigmoid(out[1]) + grid_y] + [out[2 * j + 1] + grid_y for j in range(1, K)]
    det_confs = torch.sigmoid(out[2 * K])
    cls_confs = torch.softmax(out[2 * K + 1:2 * K + 1 + nC].transpose(0, 1), dim=1)
    cls_max_confs, cls_max_ids = torch.max(cls_confs, 1)
    sz_hw, sz_hwa = h * w, h * w * nA
    all_boxes = []
    max_cls_conf = -float("inf")
    max_ind = None
    for b in range(batch):
        boxes = []
        max_conf = -1
        for cy in range(h):
            for cx in range(w):
                for i in range(nA):
                    ind = b * sz_hwa + i * sz_hw + cy * w + cx
                    det_conf = det_confs[ind]
                    conf = det_confs[ind] if only_objectness else det_confs[ind] * cls_max_confs[ind]
                    if det_confs[ind] > max_conf and cls_confs[ind, correspondingclass] > max_cls_conf:
                        max_conf = det_confs[ind]
                        max_cls_conf = cls_confs[ind, correspondingclass]
                        max_ind = ind
                    if conf > conf_thresh:
                        box = []
                        for j in range(K):
                            box.append(xs[j][ind] / w)
                            box.append(ys[j][ind] / h)
                        box += [det_conf, cls_max_confs[ind], cls_max_ids[ind]]
                        if (not only_objectness) and validation:
                            for c in range(nC):
                                tmp_conf = cls_confs[ind][c]
                                if c != cls_max_ids[ind] and det_confs[ind] * tmp_conf > conf_thresh:
                                    box += [tmp_conf, c]
                        boxes.append(box)
        if len(boxes) == 0 or correspondingclass not in [int(bx[2 * K + 2]) for bx in boxes]:
            box = []
            for j in range(K):
                box.append(xs[j][max_ind] / w)
                box.append(ys[j][max_ind] / h)
            box += [max_conf, max_cls_conf, correspondingclass]
            boxes.append(box)
        all_boxes.append(boxes)
    return all_boxes
