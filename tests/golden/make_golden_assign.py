"""Golden vectors for row f3 (Hungarian target assignment): the reference's OWN HungarianAssigner3D /
BBox3DL1Cost / normalize_bbox executed unmodified in the build container (oracle/ref_loader.
load_hungarian_assigner; mmdet's FocalLossCost, un-vendored, is the restated published formula).
Run:  python tests/golden/make_golden_assign.py   ->  tests/golden/assign.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graph_detr4d_b200 import synthetic as syn  # noqa: E402
from oracle import assign_oracle as ao, ref_loader  # noqa: E402
from tests.test_assign_oracle import make_case  # noqa: E402

L, B, Q, G, C = 6, 1, 900, 37, 10
H3D, _, _ = ref_loader.load_hungarian_assigner(ao.FocalLossCost)
assigner = H3D(cls_cost=dict(type="FocalLossCost", weight=2.0), reg_cost=dict(type="BBox3DL1Cost", weight=0.25),
               iou_cost=dict(type="IoUCost", weight=0.0), pc_range=syn.PC_RANGE)
cases = [make_case(Q, G, C, seed=10 * l) for l in range(L)]           # same cases as tests/test_assign_gpu.py
gt, labels = cases[0][2], cases[0][3]
inds, labs = [], []
for l in range(L):
    res = assigner.assign(cases[l][0], cases[l][1], gt, labels)
    inds.append(res.gt_inds.numpy())
    labs.append(res.labels.numpy())
norm = ao.normalize_bbox(gt)
cost0 = assigner.cls_cost(cases[0][1], labels) + assigner.reg_cost(cases[0][0][:, :8], norm[:, :8])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "assign.npz"),
                    bbox=np.stack([c[0].numpy() for c in cases])[:, None], cls=np.stack([c[1].numpy() for c in cases])[:, None],
                    gt=gt.numpy(), labels=labels.numpy(), inds=np.stack(inds)[:, None], out_labels=np.stack(labs)[:, None],
                    cost0=cost0.numpy())
print("wrote assign.npz")
