"""Row f4 (second half): the CPU restatement of the PE head's feature position-embedding block
(oracle/fpe_oracle.py) vs the reference's own lines executed in this container
(detr3d_head_pe.py:510-553 + SELayer + SinePositionalEncoding3D via ref_loader.load_fpe_block),
and vs the committed golden fixture (runs anywhere)."""
import os
import warnings

import numpy as np
import pytest
import torch

from graph_detr4d_b200 import synthetic as syn
from oracle import fpe_oracle, ref_loader

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fpe_block.npz")


def load_fpe_golden():
    z = np.load(GOLD)
    C, D, F_, B, T = (int(z[k]) for k in ("C", "D", "F", "B", "T"))
    shapes = [tuple(int(v) for v in s) for s in z["shapes"]]
    metas = syn.make_img_metas(B, T, img_shape=tuple(z["img"][0]), pad_shape=tuple(int(v) for v in z["pad"]))
    for m in metas:
        m["img_shape"] = [tuple(int(v) for v in z["img"][i % 6]) for i in range(6 * T)]
    L = len(shapes)
    feats = [torch.from_numpy(z[f"feat{i}_bf16"].copy()).view(torch.bfloat16).float() for i in range(L)]
    t = lambda k: torch.from_numpy(z[k].copy())
    sd = {k[3:]: t(k) for k in z.files if k.startswith("sd.")}
    outs = [t(f"out{i}") for i in range(L)]
    gouts = [torch.randn(o.shape, generator=torch.Generator().manual_seed(10 + i)) for i, o in enumerate(outs)]
    return dict(C=C, D=D, F=F_, B=B, T=T, shapes=shapes, metas=metas, feats=feats, sd=sd, outs=outs, gouts=gouts,
                grads=[t(f"grad_feat{i}") for i in range(L)], masks=[t(f"mask{i}") for i in range(L)])


def test_golden_fixture_matches_restatement():
    gd = load_fpe_golden()
    feats = [f.clone().requires_grad_(True) for f in gd["feats"]]
    outs, masks = fpe_oracle.fpe_block(gd["sd"], feats, gd["metas"], gd["D"], 1, syn.PC_RANGE, True, num_feats=gd["F"])
    sum((o * g).sum() for o, g in zip(outs, gd["gouts"])).backward()
    for l in range(len(outs)):
        assert torch.equal(masks[l], gd["masks"][l])
        assert float((outs[l] - gd["outs"][l]).abs().max()) <= 2e-5 * float(gd["outs"][l].abs().max())
        assert float((feats[l].grad - gd["grads"][l]).abs().max()) <= 2e-5 * float(gd["grads"][l].abs().max())
    assert 0.02 < float(masks[0].float().mean()) < 0.5                    # the padding masks are non-trivial
    assert float(feats[0].grad[:, 6:].abs().max()) == 0.0                 # past frames of level 0 are detached (:512-516)
    assert float(feats[1].grad[:, 6:].abs().max()) > 0.0                  # ... of level 0 ONLY


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
@pytest.mark.parametrize("B,T,with_detach", [(1, 2, True), (2, 1, True), (1, 2, False)])
def test_restatement_matches_executed_reference(B, T, with_detach):
    ref = ref_loader.load_fpe_block()
    C, D, F_ = 32, 8, 16
    shapes = [(29, 50), (15, 25), (8, 13), (4, 7)]
    head = ref.make_head(embed_dims=C, depth_num=D, pc_range=syn.PC_RANGE, num_feats=F_, with_detach=with_detach, seed=2)
    metas = syn.make_img_metas(B, T, img_shape=(225, 400, 3), pad_shape=(232, 400, 3))
    sizes = [(225, 400, 3), (232, 400, 3), (200, 390, 3), (225, 400, 3), (150, 400, 3), (225, 333, 3)]
    for b, m in enumerate(metas):
        m["img_shape"] = [sizes[(i + b) % 6] for i in range(6 * T)]
    feats = [f.requires_grad_(True) for f in syn.make_feats(B, 6 * T, C, shapes, seed=3)]
    out_r, masks_r = ref.forward_fpe(head, list(feats), metas)
    g = [torch.randn(o.shape, generator=torch.Generator().manual_seed(i)) for i, o in enumerate(out_r)]
    sum((o * gg).sum() for o, gg in zip(out_r, g)).backward()
    grads_r = [f.grad.clone() for f in feats]
    for f in feats:
        f.grad = None
    out_o, masks_o = fpe_oracle.fpe_block(head.state_dict(), feats, metas, D, 1, syn.PC_RANGE, with_detach, num_feats=F_)
    sum((o * gg).sum() for o, gg in zip(out_o, g)).backward()
    for l in range(4):
        assert torch.equal(masks_o[l], masks_r[l]) and masks_r[l].any()
        assert float((out_o[l] - out_r[l]).abs().max()) <= 1e-6 * float(out_r[l].abs().max())
        assert float((feats[l].grad - grads_r[l]).abs().max()) <= 1e-6 * float(grads_r[l].abs().max())


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_sine_embedding_and_masks_alone_match_reference_classes():
    ref = ref_loader.load_fpe_block()
    pe = ref.SinePositionalEncoding3D(num_feats=128, normalize=True, offset=-0.5)
    metas = syn.make_img_metas(1, 2)                                    # 900x1600 inside 928x1600: bottom rows masked
    masks = fpe_oracle.level_masks(1, 12, [(116, 200), (15, 25)], metas)
    assert masks[0][0, 0, 113:].all() and not masks[0][0, 0, :113].any()   # 113*8 = 904 >= 900
    for m in masks:
        assert torch.equal(fpe_oracle.sine_pe3d(m, 128, offset=-0.5), pe(m))
