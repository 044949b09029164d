"""Row f3 on the GPU: fused match-cost kernel vs the CPU oracle, and the batched one-sync
assignment vs the reference's per-layer, per-sample procedure (oracle restatement)."""
import pytest
import torch

from graph_detr4d_b200.assign import BatchedHungarianAssigner3D
from oracle import assign_oracle as ao
from tests.test_assign_oracle import make_case

pytestmark = pytest.mark.gpu


def _batch(L, B, Q, Gs, C=10, seed=0):
    cases = [[make_case(Q, Gs[b], C, seed=seed + 10 * l + b) for b in range(B)] for l in range(L)]
    bbox = torch.stack([torch.stack([cases[l][b][0] for b in range(B)]) for l in range(L)])     # (L,B,Q,10)
    cls = torch.stack([torch.stack([cases[l][b][1] for b in range(B)]) for l in range(L)])
    gts = [cases[0][b][2] for b in range(B)]                                                   # gt is per sample
    labs = [cases[0][b][3] for b in range(B)]
    return bbox, cls, gts, labs


@pytest.mark.parametrize("L,B,Q,Gs", [(6, 1, 900, [37]), (3, 2, 120, [5, 130]), (2, 3, 64, [0, 7, 1])])
def test_batched_assignment_matches_per_layer_reference(L, B, Q, Gs):
    bbox, cls, gts, labs = _batch(L, B, Q, Gs)
    asg = BatchedHungarianAssigner3D()
    buf, layout = asg.match_costs(bbox.cuda(), cls.cuda(), [g.cuda() for g in gts], [l.cuda() for l in labs])
    for b, (off, G) in enumerate(layout):
        if G == 0:
            continue
        got = buf[off:off + L * Q * G].view(L, Q, G).cpu()
        for l in range(L):
            want = ao.match_cost(bbox[l, b], cls[l, b], gts[b], labs[b])
            assert float((got[l] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    inds, labels = asg.assign_layers(bbox.cuda(), cls.cuda(), [g.cuda() for g in gts], [l.cuda() for l in labs])
    assert tuple(inds.shape) == (L, B, Q) and inds.dtype == torch.long
    for l in range(L):
        for b in range(B):
            wi, wl = ao.hungarian_assign(bbox[l, b], cls[l, b], gts[b], labs[b])
            assert torch.equal(inds[l, b].cpu(), wi) and torch.equal(labels[l, b].cpu(), wl)


def test_nan_and_inf_predictions_are_sanitised():
    bbox, cls, gts, labs = _batch(2, 1, 40, [5], seed=5)
    bbox[0, 0, 3, 2] = float("nan")
    cls[1, 0, 7, :] = float("inf")
    asg = BatchedHungarianAssigner3D()
    buf, _ = asg.match_costs(bbox.cuda(), cls.cuda(), [gts[0].cuda()], [labs[0].cuda()])
    assert bool(torch.isfinite(buf).all())
    inds, _ = asg.assign_layers(bbox.cuda(), cls.cuda(), [gts[0].cuda()], [labs[0].cuda()])
    for l in range(2):
        wi, _ = ao.hungarian_assign(bbox[l, 0], cls[l, 0], gts[0], labs[0])
        assert torch.equal(inds[l, 0].cpu(), wi)


def test_assigner_refuses_cpu_tensors():
    bbox, cls, gts, labs = _batch(1, 1, 8, [2])
    with pytest.raises(RuntimeError):
        BatchedHungarianAssigner3D().assign_layers(bbox, cls, gts, labs)


def test_batched_assignment_matches_reference_golden():
    """Same inputs as the [6-1-900] case above, expected outputs frozen from the reference's own
    HungarianAssigner3D executed in the build container (tests/golden/make_golden_assign.py)."""
    import os
    import numpy as np
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assign.npz"))
    bbox, cls = torch.as_tensor(gd["bbox"]).cuda(), torch.as_tensor(gd["cls"]).cuda()
    gt, labels = torch.as_tensor(gd["gt"]).cuda(), torch.as_tensor(gd["labels"]).cuda()
    asg = BatchedHungarianAssigner3D()
    inds, labs = asg.assign_layers(bbox, cls, [gt], [labels])
    assert torch.equal(inds.cpu(), torch.as_tensor(gd["inds"])) and torch.equal(labs.cpu(), torch.as_tensor(gd["out_labels"]))
    buf, layout = asg.match_costs(bbox, cls, [gt], [labels])
    got = buf[:900 * 37].view(900, 37).cpu()
    want = torch.as_tensor(gd["cost0"])
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
