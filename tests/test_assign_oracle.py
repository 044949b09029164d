"""Row f3: the CPU restatement of the Hungarian target assignment vs the reference's own
HungarianAssigner3D / BBox3DL1Cost / normalize_bbox executed unmodified in this container."""
import pytest
import torch

from graph_detr4d_b200 import synthetic as syn
from oracle import assign_oracle as ao, ref_loader


def make_case(Q=900, G=37, C=10, seed=0, code=10):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(Q, C, generator=g) * 2 - 2
    bbox = torch.randn(Q, code, generator=g)
    gt = torch.randn(G, 9, generator=g)
    gt[:, 3:6] = torch.rand(G, 3, generator=g) * 4 + 0.3            # sizes > 0 (log)
    labels = torch.randint(0, C, (G,), generator=g)
    return bbox, cls, gt, labels


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
@pytest.mark.parametrize("Q,G,seed", [(900, 37, 0), (300, 1, 1), (50, 80, 2), (64, 0, 3)])
def test_restatement_matches_executed_reference(Q, G, seed):
    H3D, L1, norm = ref_loader.load_hungarian_assigner(ao.FocalLossCost)
    assigner = H3D(cls_cost=dict(type="FocalLossCost", weight=2.0), reg_cost=dict(type="BBox3DL1Cost", weight=0.25),
                   iou_cost=dict(type="IoUCost", weight=0.0), pc_range=syn.PC_RANGE)
    bbox, cls, gt, labels = make_case(Q, G, seed=seed)
    res = assigner.assign(bbox, cls, gt, labels)
    inds, lab = ao.hungarian_assign(bbox, cls, gt, labels)
    assert res.num_gts == G and torch.equal(res.gt_inds, inds) and torch.equal(res.labels, lab)
    if G:
        assert torch.equal(norm(gt, syn.PC_RANGE), ao.normalize_bbox(gt))
        want = assigner.cls_cost(cls, labels) + assigner.reg_cost(bbox[:, :8], norm(gt, syn.PC_RANGE)[:, :8])
        assert torch.equal(torch.nan_to_num(want, nan=100.0, posinf=100.0, neginf=-100.0),
                           ao.match_cost(bbox, cls, gt, labels))
        assert int((inds > 0).sum()) == min(Q, G)


def test_nan_predictions_are_sanitised_like_the_reference():
    bbox, cls, gt, labels = make_case(40, 5, seed=5)
    bbox[3, 2] = float("nan")
    cls[7, :] = float("inf")
    cost = ao.match_cost(bbox, cls, gt, labels)
    assert torch.isfinite(cost).all() and float(cost[3].max()) == 100.0
    inds, _ = ao.hungarian_assign(bbox, cls, gt, labels)
    assert int((inds > 0).sum()) == 5


def test_threaded_solver_equals_sequential():
    """The product's thread-pool solve (host logic, no GPU) returns exactly scipy's sequential results."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    from graph_detr4d_b200.assign import solve_all
    rng = np.random.default_rng(0)
    mats = [rng.standard_normal((200, g)).astype(np.float32) for g in (40, 1, 13, 250, 40, 7)]
    for (r1, c1), m in zip(solve_all(mats), mats):
        r2, c2 = linear_sum_assignment(m)
        assert np.array_equal(r1, r2) and np.array_equal(c1, c2)


def test_oracle_reproduces_reference_golden():
    """tests/golden/assign.npz = the reference's own HungarianAssigner3D executed in the build container."""
    import os
    import numpy as np
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assign.npz"))
    gt, labels = torch.as_tensor(gd["gt"]), torch.as_tensor(gd["labels"])
    for l in range(gd["bbox"].shape[0]):
        bbox, cls = torch.as_tensor(gd["bbox"][l, 0]), torch.as_tensor(gd["cls"][l, 0])
        inds, lab = ao.hungarian_assign(bbox, cls, gt, labels)
        assert torch.equal(inds, torch.as_tensor(gd["inds"][l, 0])) and torch.equal(lab, torch.as_tensor(gd["out_labels"][l, 0]))
    assert torch.equal(ao.match_cost(torch.as_tensor(gd["bbox"][0, 0]), torch.as_tensor(gd["cls"][0, 0]), gt, labels),
                       torch.as_tensor(gd["cost0"]))
