"""Row f3 timing: batched one-sync assignment (graph_detr4d_b200.assign) vs the reference-style loop
(per layer: ~15 torch ops for the cost matrix, .cpu() sync, scipy, indices back), 6 layers, B=1, Q=900,
G=40 ground-truth boxes.  Wall clock around a synchronize (host syncs are the point).  (dev tool)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy.optimize import linear_sum_assignment
from graph_detr4d_b200.assign import BatchedHungarianAssigner3D

L, B, Q, G, C = 6, 1, 900, 40, 10
g = torch.Generator().manual_seed(0)
bbox = torch.randn(L, B, Q, 10, generator=g).cuda(); cls = (torch.randn(L, B, Q, C, generator=g) * 2 - 2).cuda()
gt = torch.randn(G, 9, generator=g); gt[:, 3:6] = torch.rand(G, 3, generator=g) * 4 + 0.3
gt = gt.cuda(); lab = torch.randint(0, C, (G,), generator=g).cuda()


def per_layer():                                   # the reference's procedure, restated with torch ops
    out = []
    for l in range(L):
        p = cls[l, 0].sigmoid()
        neg = -(1 - p + 1e-12).log() * 0.75 * p.pow(2)
        pos = -(p + 1e-12).log() * 0.25 * (1 - p).pow(2)
        cc = (pos[:, lab] - neg[:, lab]) * 2.0
        n = torch.cat([gt[:, 0:1], gt[:, 1:2], gt[:, 3:4].log(), gt[:, 4:5].log(), gt[:, 2:3], gt[:, 5:6].log(),
                       gt[:, 6:7].sin(), gt[:, 6:7].cos()], -1)
        cost = cc + torch.cdist(bbox[l, 0, :, :8], n, p=1) * 0.25
        cost = torch.nan_to_num(cost.detach().cpu(), nan=100.0, posinf=100.0, neginf=-100.0)     # sync
        r, c = linear_sum_assignment(cost)
        inds = torch.zeros(Q, dtype=torch.long, device="cuda")
        inds[torch.from_numpy(r).cuda()] = torch.from_numpy(c).cuda() + 1
        out.append(inds)
    return out


asg = BatchedHungarianAssigner3D()
batched = lambda: asg.assign_layers(bbox, cls, [gt], [lab])


def wall(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


a, b = wall(per_layer), wall(batched)
ref = torch.stack(per_layer()); got = batched()[0][:, 0]
print(json.dumps(dict(layers=L, queries=Q, gts=G, per_layer_ms=a, batched_ms=b, speedup=a / b,
                      identical=bool(torch.equal(ref, got)), host_syncs=dict(per_layer=L, batched=1))))
