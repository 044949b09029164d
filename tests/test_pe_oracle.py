"""Row f4: the CPU restatement of Detr3DHeadPE.position_embeding's frustum input vs the
unmodified reference method executed in this container, and vs the committed golden fixture."""
import os
import types

import numpy as np
import pytest
import torch

from graph_detr4d_b200 import synthetic as syn
from oracle import pe_oracle, ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frustum_pe.npz")
SHAPES = [(6, 10), (3, 5), (2, 3), (1, 2)]
D, DEPTH_START = 8, 1


def _masks(B, N, shapes, seed=3):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(B, N, H, W, generator=g) > 0.8 for H, W in shapes]


def _reference(B, T, shapes, depth_num, masks):
    fn = ref_loader.load_position_embeding()
    metas = syn.make_img_metas(B, T)
    stub = types.SimpleNamespace(depth_num=depth_num, depth_start=DEPTH_START, pc_range=syn.PC_RANGE,
                                 position_encoder=lambda x: x, embed_dims=3 * depth_num)
    feats = [torch.zeros(B, 6 * T, 4, H, W) for H, W in shapes]
    with torch.no_grad():
        emb, m = fn(stub, feats, metas, masks)
    # the stub encoder is the identity, so `emb` is the conv INPUT viewed (B,N,embed_dims=3D,H,W)
    return [e.flatten(0, 1) for e in emb], m, metas


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
@pytest.mark.parametrize("B,T,depth_num", [(1, 1, 8), (2, 2, 4), (1, 2, 64)])
def test_restatement_matches_executed_reference(B, T, depth_num):
    masks = _masks(B, 6 * T, SHAPES)
    xr, mr, metas = _reference(B, T, SHAPES, depth_num, masks)
    xo, mo = pe_oracle.frustum_pe_input(SHAPES, metas, depth_num, DEPTH_START, syn.PC_RANGE, masks)
    for a, b, ma, mb in zip(xo, xr, mo, mr):
        assert a.shape == b.shape and ma.shape == mb.shape
        assert torch.equal(ma, mb)                                         # mask bit-exact
        # compared in the normalised-coordinate domain: inverse_sigmoid amplifies 1-ulp differences of
        # the 4x4 mat-vec (BLAS order in the reference) by 1/(x(1-x)) <= 1e5 next to the clamp
        assert float((a.sigmoid() - b.sigmoid()).abs().max()) <= 2e-6
        ok = (b.sigmoid() > 1e-3) & (b.sigmoid() < 1 - 1e-3)
        assert float((a - b)[ok].abs().max()) <= 2e-3


def test_golden_fixture_matches_restatement():
    gd = np.load(GOLD, allow_pickle=False)
    B, T, depth_num = int(gd["B"]), int(gd["T"]), int(gd["depth_num"])
    shapes = [tuple(s) for s in gd["shapes"]]
    metas = syn.make_img_metas(B, T)
    masks = [torch.as_tensor(gd[f"mask_in{l}"]) for l in range(len(shapes))]
    xo, mo = pe_oracle.frustum_pe_input(shapes, metas, depth_num, DEPTH_START, syn.PC_RANGE, masks)
    for l in range(len(shapes)):
        assert torch.equal(mo[l], torch.as_tensor(gd[f"mask{l}"]))
        assert float((xo[l].sigmoid() - torch.as_tensor(gd[f"x{l}"]).sigmoid()).abs().max()) <= 2e-6
