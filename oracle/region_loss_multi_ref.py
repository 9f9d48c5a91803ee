"""Oracle: multi-object RegionLoss (reference multi_obj_pose_estimation/region_loss_multi.py:9-189,
utils_multi.py:125-156 bbox_iou) restated on CPU tensors and differentiated by autograd.

Quirk reproduced on purpose: ``pred_box`` is read with ``best_n = -1`` BEFORE the anchor is chosen
(region_loss_multi.py:51,63), i.e. at flat index ``b*nAnchors - nPixels + cell``: the LAST anchor of the PREVIOUS
image (wrapping to the last image for b = 0).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import torch

from .region_loss_ref import corner_confidences_ref, corner_confidence_ref


def bbox_iou_ref(box1, box2):
    """utils_multi.py:125-156 with x1y1x2y2=False (centre format)."""
    mx = min(box1[0] - box1[2] / 2.0, box2[0] - box2[2] / 2.0)
    Mx = max(box1[0] + box1[2] / 2.0, box2[0] + box2[2] / 2.0)
    my = min(box1[1] - box1[3] / 2.0, box2[1] - box2[3] / 2.0)
    My = max(box1[1] + box1[3] / 2.0, box2[1] + box2[3] / 2.0)
    w1, h1, w2, h2 = box1[2], box1[3], box2[2], box2[3]
    cw, ch = w1 + w2 - (Mx - mx), h1 + h2 - (My - my)
    if cw <= 0 or ch <= 0:
        return 0.0
    carea = cw * ch
    return carea / (w1 * h1 + w2 * h2 - carea)


def build_targets_multi_ref(pred_corners, target, num_keypoints, anchors, num_anchors, num_classes, nH, nW,
                            noobject_scale, object_scale, sil_thresh):
    nB, nA, K = target.size(0), num_anchors, num_keypoints
    anchor_step = len(anchors) // num_anchors
    conf_mask = torch.ones(nB, nA, nH, nW) * noobject_scale
    coord_mask = torch.zeros(nB, nA, nH, nW)
    cls_mask = torch.zeros(nB, nA, nH, nW)
    txs = [torch.zeros(nB, nA, nH, nW) for _ in range(K)]
    tys = [torch.zeros(nB, nA, nH, nW) for _ in range(K)]
    tconf = torch.zeros(nB, nA, nH, nW)
    tcls = torch.zeros(nB, nA, nH, nW)
    nl = 2 * K + 3
    nAnchors, nPixels = nA * nH * nW, nH * nW
    for b in range(nB):                                                    # :28-41
        cur_pred = pred_corners[b * nAnchors:(b + 1) * nAnchors].t()
        cur_confs = torch.zeros(nAnchors)
        for t in range(50):
            if target[b][t * nl + 1] == 0:
                break
            g = [float(target[b][t * nl + 1 + j]) for j in range(2 * K)]
            gt = torch.FloatTensor(g).repeat(nAnchors, 1).t()
            cur_confs = torch.max(cur_confs, corner_confidences_ref(cur_pred, gt))
        conf_mask[b][cur_confs.view_as(conf_mask[b]) > sil_thresh] = 0
    nGT = nCorrect = 0
    for b in range(nB):                                                    # :45-90
        for t in range(50):
            if target[b][t * nl + 1] == 0:
                break
            nGT += 1
            best_iou, best_n = 0.0, -1
            gt_box = [target[b][t * nl + 1 + j] for j in range(2 * K)]
            gx = [target[b][t * nl + 2 * i + 1] * nW for i in range(K)]
            gy = [target[b][t * nl + 2 * i + 2] * nH for i in range(K)]
            gi0, gj0 = int(gx[0]), int(gy[0])
            pred_box = pred_corners[b * nAnchors + best_n * nPixels + gj0 * nW + gi0]     # best_n == -1 here (reference bug)
            conf = corner_confidence_ref(gt_box, pred_box)
            gw = target[b][t * nl + nl - 2] * nW
            gh = target[b][t * nl + nl - 1] * nH
            for n in range(nA):
                iou = bbox_iou_ref([0, 0, anchors[anchor_step * n], anchors[anchor_step * n + 1]], [0, 0, gw, gh])
                if iou > best_iou:
                    best_iou, best_n = iou, n
            coord_mask[b][best_n][gj0][gi0] = 1
            cls_mask[b][best_n][gj0][gi0] = 1
            conf_mask[b][best_n][gj0][gi0] = object_scale
            for i in range(K):
                txs[i][b][best_n][gj0][gi0] = gx[i] - gi0
                tys[i][b][best_n][gj0][gi0] = gy[i] - gj0
            tconf[b][best_n][gj0][gi0] = conf
            tcls[b][best_n][gj0][gi0] = target[b][t * nl]
            if conf > 0.5:
                nCorrect += 1
    return nGT, nCorrect, coord_mask, conf_mask, cls_mask, txs, tys, tconf, tcls


def region_loss_multi_ref(output, target, epoch, anchors, num_keypoints=9, num_classes=13, num_anchors=5,
                          coord_scale=1.0, noobject_scale=1.0, object_scale=5.0, class_scale=1.0, thresh=0.6,
                          pretrain_num_epochs=15, build_targets=build_targets_multi_ref):
    """region_loss_multi.py:110-189 on CPU tensors -> (loss, info)."""
    nB, nA, nC, K = output.size(0), num_anchors, num_classes, num_keypoints
    nH, nW = output.size(2), output.size(3)
    out = output.view(nB, nA, 2 * K + 1 + nC, nH, nW)
    x = [torch.sigmoid(out[:, :, 0])] + [out[:, :, 2 * i] for i in range(1, K)]
    y = [torch.sigmoid(out[:, :, 1])] + [out[:, :, 2 * i + 1] for i in range(1, K)]
    conf = torch.sigmoid(out[:, :, 2 * K])
    cls = out[:, :, 2 * K + 1:2 * K + 1 + nC]
    cls = cls.contiguous().view(nB * nA, nC, nH * nW).transpose(1, 2).contiguous().view(nB * nA * nH * nW, nC)
    N = nB * nA * nH * nW
    grid_x = torch.linspace(0, nW - 1, nW).repeat(nH, 1).repeat(nB * nA, 1, 1).view(N)
    grid_y = torch.linspace(0, nH - 1, nH).repeat(nW, 1).t().repeat(nB * nA, 1, 1).view(N)
    pc = torch.zeros(2 * K, N)
    for i in range(K):
        pc[2 * i] = (x[i].detach().reshape(N) + grid_x) / nW
        pc[2 * i + 1] = (y[i].detach().reshape(N) + grid_y) / nH
    pred_corners = pc.t().contiguous().view(-1, 2 * K)
    nGT, nCorrect, coord_mask, conf_mask, cls_mask, txs, tys, tconf, tcls = build_targets(
        pred_corners, target.detach().float(), K, anchors, nA, nC, nH, nW, noobject_scale, object_scale, thresh)
    cls_mask_b = cls_mask == 1
    nProposals = int((conf > 0.25).sum())
    tcls_sel = tcls[cls_mask_b].long()
    conf_mask = conf_mask.sqrt()
    cls_sel = cls[cls_mask_b.view(-1, 1).repeat(1, nC)].view(-1, nC)
    loss_x = sum(coord_scale * ((x[i] * coord_mask - txs[i] * coord_mask) ** 2).sum() / 2.0 for i in range(K))
    loss_y = sum(coord_scale * ((y[i] * coord_mask - tys[i] * coord_mask) ** 2).sum() / 2.0 for i in range(K))
    loss_conf = ((conf * conf_mask - tconf * conf_mask) ** 2).sum() / 2.0
    loss_cls = class_scale * torch.nn.functional.cross_entropy(cls_sel, tcls_sel, reduction="sum")
    loss = loss_x + loss_y + loss_cls + (loss_conf if epoch > pretrain_num_epochs else 0.0)
    return loss, dict(loss_x=loss_x, loss_y=loss_y, loss_conf=loss_conf, loss_cls=loss_cls, nGT=nGT, nCorrect=nCorrect,
                      nProposals=nProposals, tconf=tconf, conf_mask=conf_mask, coord_mask=coord_mask, tcls=tcls)
