"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Hungarian target assignment of the
reference's loss:

    HungarianAssigner3D.assign     core/bbox/assigners/hungarian_assigner_3d.py:60-145
    BBox3DL1Cost                   core/bbox/match_costs/match_cost.py:6-28   (torch.cdist, p=1)
    normalize_bbox                 core/bbox/util.py:38-57
    FocalLossCost                  mmdet 2.x core/bbox/match_costs/match_cost.py -- third-party,
                                   NOT vendored, version un-pinned by the reference (used at
                                   projects/configs/detr3d/detr3d_res50.py:112); its published
                                   formula is restated below.

Pinned: tests/test_assign_oracle.py executes the reference's own HungarianAssigner3D /
BBox3DL1Cost / normalize_bbox unmodified (oracle/ref_loader.load_hungarian_assigner) with this
FocalLossCost plugged in and finds identical assignments and costs.
"""
from __future__ import annotations

import torch
from scipy.optimize import linear_sum_assignment


class FocalLossCost:
    """mmdet 2.x FocalLossCost (weight, alpha=0.25, gamma=2, eps=1e-12)."""

    def __init__(self, weight=1.0, alpha=0.25, gamma=2, eps=1e-12):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

    def __call__(self, cls_pred, gt_labels):
        cls_pred = cls_pred.sigmoid()
        neg_cost = -(1 - cls_pred + self.eps).log() * (1 - self.alpha) * cls_pred.pow(self.gamma)
        pos_cost = -(cls_pred + self.eps).log() * self.alpha * (1 - cls_pred).pow(self.gamma)
        cls_cost = pos_cost[:, gt_labels] - neg_cost[:, gt_labels]
        return cls_cost * self.weight


def normalize_bbox(bboxes):
    """core/bbox/util.py:38-57: (cx, cy, cz, w, l, h, rot, vx, vy) -> (cx, cy, log w, log l, cz, log h, sin, cos, vx, vy)."""
    cx, cy, cz = bboxes[..., 0:1], bboxes[..., 1:2], bboxes[..., 2:3]
    w, l, h = bboxes[..., 3:4].log(), bboxes[..., 4:5].log(), bboxes[..., 5:6].log()
    rot = bboxes[..., 6:7]
    parts = [cx, cy, w, l, cz, h, rot.sin(), rot.cos()]
    if bboxes.size(-1) > 7:
        parts += [bboxes[..., 7:8], bboxes[..., 8:9]]
    return torch.cat(parts, dim=-1)


def match_cost(bbox_pred, cls_pred, gt_bboxes, gt_labels, cls_weight=2.0, reg_weight=0.25):
    """:117-131 -- weighted cost matrix (num_query, num_gt) after nan_to_num."""
    cls_cost = FocalLossCost(cls_weight)(cls_pred, gt_labels)                         # :118
    ngt = normalize_bbox(gt_bboxes)                                                   # :120
    reg_cost = torch.cdist(bbox_pred[:, :8], ngt[:, :8], p=1) * reg_weight            # :121
    cost = cls_cost + reg_cost                                                        # :124
    return torch.nan_to_num(cost.detach().cpu(), nan=100.0, posinf=100.0, neginf=-100.0)   # :127-131


def hungarian_assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, cls_weight=2.0, reg_weight=0.25):
    """-> (assigned_gt_inds (Q,) long: 0 background / 1-based gt index, assigned_labels (Q,) long: -1 / label)."""
    num_gts, num_bboxes = gt_bboxes.size(0), bbox_pred.size(0)
    inds = torch.full((num_bboxes,), -1, dtype=torch.long)
    labels = torch.full((num_bboxes,), -1, dtype=torch.long)
    if num_gts == 0 or num_bboxes == 0:                                               # :104-110
        if num_gts == 0:
            inds[:] = 0
        return inds, labels
    cost = match_cost(bbox_pred, cls_pred, gt_bboxes, gt_labels, cls_weight, reg_weight)
    rows, cols = linear_sum_assignment(cost)                                          # :132
    rows, cols = torch.from_numpy(rows), torch.from_numpy(cols)
    inds[:] = 0                                                                       # :140
    inds[rows] = cols + 1                                                             # :142
    labels[rows] = gt_labels[cols]                                                    # :143
    return inds, labels
