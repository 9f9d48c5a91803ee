"""Oracle: single-object RegionLoss (reference region_loss.py:9-175, utils.py:138-187).

The reference's ``RegionLoss.forward`` cannot run on CPU (hard-coded ``torch.cuda``) nor on
modern torch (0-dim ``.data[0]``), so it is restated here line by line on CPU tensors and
differentiated by autograd.  ``build_targets`` / ``corner_confidence(s)`` are restated too and
pinned against the reference's own functions by tests/golden/make_golden.py.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import torch


def corner_confidences_ref(gt_corners, pr_corners, th=80, sharpness=2, im_width=640, im_height=480):
    """utils.py:138-165.  (2K x nA) tensors -> (nA,) mean keypoint confidence (no eps)."""
    nA = gt_corners.shape[1]
    dist = (gt_corners - pr_corners).t().contiguous().view(nA, -1, 2).clone()
    dist[:, :, 0] *= im_width
    dist[:, :, 1] *= im_height
    d = torch.sqrt((dist ** 2).sum(2))
    mask = (d < th).float()
    conf = torch.exp(sharpness * (1 - d / th)) - 1
    conf0 = (torch.exp(torch.tensor([float(sharpness)])) - 1).item()   # exp(2*(1-0)) - 1 in fp32
    conf = mask * (conf / conf0)
    return conf.mean(1)


def corner_confidence_ref(gt_corners, pr_corners, th=80, sharpness=2, im_width=640, im_height=480):
    """utils.py:167-187.  two length-2K vectors -> scalar (eps=1e-5 in the denominator)."""
    dist = (torch.as_tensor(gt_corners, dtype=torch.float32) - pr_corners).view(-1, 2).clone()
    dist[:, 0] *= im_width
    dist[:, 1] *= im_height
    d = torch.sqrt((dist ** 2).sum(1))
    mask = (d < th).float()
    conf = torch.exp(sharpness * (1.0 - d / th)) - 1
    conf0 = torch.exp(torch.tensor([float(sharpness)])) - 1 + 1e-5
    # the reference divides a (K,) vector by a (K,1) tensor: a (K,K) broadcast whose mean equals the
    # mean over keypoints (utils.py:184-187); kept so that fp32 summation order matches bit for bit
    K = d.numel()
    return (mask * (conf / conf0.repeat(K, 1))).mean()


def build_targets_ref(pred_corners, target, num_keypoints, num_anchors, num_classes, nH, nW,
                      noobject_scale, object_scale, sil_thresh):
    """region_loss.py:9-78 (single-object: exactly one ground truth per image)."""
    nB, nA, K = target.size(0), num_anchors, num_keypoints
    conf_mask = torch.ones(nB, nA, nH, nW) * noobject_scale
    coord_mask = torch.zeros(nB, nA, nH, nW)
    cls_mask = torch.zeros(nB, nA, nH, nW)
    txs = [torch.zeros(nB, nA, nH, nW) for _ in range(K)]
    tys = [torch.zeros(nB, nA, nH, nW) for _ in range(K)]
    tconf = torch.zeros(nB, nA, nH, nW)
    tcls = torch.zeros(nB, nA, nH, nW)
    num_labels = 2 * K + 3
    nAnchors, nPixels = nA * nH * nW, nH * nW
    for b in range(nB):                                            # :27-40
        cur_pred = pred_corners[b * nAnchors:(b + 1) * nAnchors].t()
        cur_confs = torch.zeros(nAnchors)
        for t in range(50):
            if target[b][t * num_labels + 1] == 0:
                break
            g = [float(target[b][t * num_labels + 1 + j]) for j in range(2 * K)]
            gt = torch.FloatTensor(g).repeat(nAnchors, 1).t()
            cur_confs = torch.max(cur_confs, corner_confidences_ref(cur_pred, gt))
        conf_mask[b][cur_confs.view_as(conf_mask[b]) > sil_thresh] = 0
    nGT = nCorrect = 0
    for b in range(nB):                                            # :45-76
        for t in range(50):
            if target[b][t * num_labels + 1] == 0:
                break
            nGT += 1
            gt_box = [target[b][t * num_labels + 1 + j] for j in range(2 * K)]
            gx = [target[b][t * num_labels + 2 * i + 1] * nW for i in range(K)]
            gy = [target[b][t * num_labels + 2 * i + 2] * nH for i in range(K)]
            gi0, gj0 = int(gx[0]), int(gy[0])
            best_n = 0
            pred_box = pred_corners[b * nAnchors + best_n * nPixels + gj0 * nW + gi0]
            conf = corner_confidence_ref(gt_box, pred_box)
            coord_mask[b][best_n][gj0][gi0] = 1
            cls_mask[b][best_n][gj0][gi0] = 1
            conf_mask[b][best_n][gj0][gi0] = object_scale
            for i in range(K):
                txs[i][b][best_n][gj0][gi0] = gx[i] - gi0
                tys[i][b][best_n][gj0][gi0] = gy[i] - gj0
            tconf[b][best_n][gj0][gi0] = conf
            tcls[b][best_n][gj0][gi0] = target[b][t * num_labels]
            if conf > 0.5:
                nCorrect += 1
    return nGT, nCorrect, coord_mask, conf_mask, cls_mask, txs, tys, tconf, tcls


def region_loss_ref(output, target, epoch, num_keypoints=9, num_classes=1, num_anchors=1,
                    coord_scale=1.0, noobject_scale=1.0, object_scale=5.0, thresh=0.6,
                    pretrain_num_epochs=15, build_targets=build_targets_ref):
    """region_loss.py:95-175 on CPU tensors.  Returns (loss, dict(parts and counters)).
    ``output`` may require grad; ``loss.backward()`` then yields the gradient oracle."""
    nB, nA, nC, K = output.size(0), num_anchors, num_classes, num_keypoints
    nH, nW = output.size(2), output.size(3)
    out = output.view(nB, nA, 2 * K + 1 + nC, nH, nW)
    x = [torch.sigmoid(out[:, :, 0])] + [out[:, :, 2 * i] for i in range(1, K)]          # :106-113
    y = [torch.sigmoid(out[:, :, 1])] + [out[:, :, 2 * i + 1] for i in range(1, K)]
    conf = torch.sigmoid(out[:, :, 2 * K])                                                # :114
    N = nB * nA * nH * nW
    grid_x = torch.linspace(0, nW - 1, nW).repeat(nH, 1).repeat(nB * nA, 1, 1).view(N)    # :121-122
    grid_y = torch.linspace(0, nH - 1, nH).repeat(nW, 1).t().repeat(nB * nA, 1, 1).view(N)
    pc = torch.zeros(2 * K, N)
    for i in range(K):                                                                    # :123-125
        pc[2 * i] = (x[i].detach().reshape(N) + grid_x) / nW
        pc[2 * i + 1] = (y[i].detach().reshape(N) + grid_y) / nH
    pred_corners = pc.t().contiguous().view(-1, 2 * K)                                    # :126-127
    nGT, nCorrect, coord_mask, conf_mask, cls_mask, txs, tys, tconf, tcls = build_targets(
        pred_corners, target.detach().float(), K, nA, nC, nH, nW, noobject_scale, object_scale, thresh)
    nProposals = int((conf > 0.25).sum())                                                 # :134
    conf_mask = conf_mask.sqrt()                                                          # :141
    loss_x = sum(coord_scale * ((x[i] * coord_mask - txs[i] * coord_mask) ** 2).sum() / 2.0 for i in range(K))
    loss_y = sum(coord_scale * ((y[i] * coord_mask - tys[i] * coord_mask) ** 2).sum() / 2.0 for i in range(K))
    loss_conf = ((conf * conf_mask - tconf * conf_mask) ** 2).sum() / 2.0                 # :152
    loss = loss_x + loss_y + loss_conf if epoch > pretrain_num_epochs else loss_x + loss_y  # :156-161
    return loss, dict(loss_x=loss_x, loss_y=loss_y, loss_conf=loss_conf, nGT=nGT, nCorrect=nCorrect,
                      nProposals=nProposals, tconf=tconf, conf_mask=conf_mask, coord_mask=coord_mask)
