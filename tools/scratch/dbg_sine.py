import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from graph_detr4d_b200 import fpe, synthetic as syn
from oracle import fpe_oracle
from tests import helpers as H
B, T = 1, 2
metas = syn.make_img_metas(B, T); N = 6 * T
wm = fpe_oracle.level_masks(B, N, H.FULL_SHAPES, metas)
for l, (h, w) in enumerate(H.FULL_SHAPES):
    got = fpe.sine_pe3d((h, w), metas, N, 128, offset=-0.5).cpu().view(B, N, 384, h, w)
    want = fpe_oracle.sine_pe3d(wm[l], 128, offset=-0.5)
    wantg = fpe_oracle.sine_pe3d(wm[l].cuda(), 128, offset=-0.5).cpu() if False else None
    d = (got - want).abs()
    idx = torch.nonzero(d == d.max())[0].tolist()
    print(l, float(d.max()), idx, float(got[tuple(idx)]), float(want[tuple(idx)]), int((d > 2e-6).sum()), d.numel())
    bad = torch.nonzero(d > 2e-6)
    if len(bad):
        print(" chan hist", torch.bincount(bad[:, 2], minlength=384).nonzero().flatten()[:20].tolist(), "rows", bad[:, 3].unique()[:10].tolist(), "cams", bad[:,1].unique().tolist())
