"""Row f4 on the GPU: the fused frustum position-embedding input kernel vs the CPU oracle
(bit-exact mask; logits equal up to logf's last ulp), vs the committed reference golden
fixture, at the nuScenes size (D = 64, 928x1600, 12 cameras), and the drop-in wrapper."""
import os

import numpy as np
import pytest
import torch

from graph_detr4d_b200 import frustum, synthetic as syn
from oracle import pe_oracle
from tests.test_pe_oracle import DEPTH_START, GOLD, SHAPES, _masks

pytestmark = pytest.mark.gpu


def _check(xs, ms, xo, mo):
    for a, b, ma, mb in zip(xs, xo, ms, mo):
        assert a.shape == b.shape and ma.dtype == torch.bool
        assert torch.equal(ma.cpu(), mb)                                           # mask bit-exact
        d = (a.cpu() - b).abs()
        assert float(d.max()) <= 1e-5 * float(b.abs().max())                       # fp32: logf ulp only


@pytest.mark.parametrize("B,T,depth_num,with_masks", [(1, 1, 8, True), (2, 2, 64, False), (1, 2, 5, True)])
def test_frustum_kernel_matches_oracle(B, T, depth_num, with_masks):
    metas = syn.make_img_metas(B, T)
    masks = _masks(B, 6 * T, SHAPES) if with_masks else None
    xo, mo = pe_oracle.frustum_pe_input(SHAPES, metas, depth_num, DEPTH_START, syn.PC_RANGE, masks)
    xs, ms = frustum.frustum_position_input(SHAPES, metas, depth_num, DEPTH_START, syn.PC_RANGE,
                                            None if masks is None else [m.cuda() for m in masks])
    _check(xs, ms, xo, mo)


def test_frustum_kernel_matches_reference_golden():
    gd = np.load(GOLD, allow_pickle=False)
    B, T, depth_num = int(gd["B"]), int(gd["T"]), int(gd["depth_num"])
    shapes = [tuple(int(v) for v in s) for s in gd["shapes"]]
    masks = [torch.as_tensor(gd[f"mask_in{l}"]).cuda() for l in range(len(shapes))]
    xs, ms = frustum.frustum_position_input(shapes, syn.make_img_metas(B, T), depth_num, DEPTH_START,
                                            syn.PC_RANGE, masks)
    for l in range(len(shapes)):
        assert torch.equal(ms[l].cpu(), torch.as_tensor(gd[f"mask{l}"]))
        ref = torch.as_tensor(gd[f"x{l}"])
        assert float((xs[l].cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_frustum_full_size_and_drop_in_wrapper():
    """nuScenes size: level 0 of 928x1600 (116x200), 12 cameras, D = 64 -> 192-channel conv input.
    Checked against the oracle on a strided subset of pixels (the kernel is per-pixel independent)
    and through the reference-shaped wrapper with a real position_encoder."""
    metas = syn.make_img_metas(1, 2)
    shapes = syn.LEVEL_SHAPES_928x1600[:2]
    xs, ms = frustum.frustum_position_input(shapes, metas, 64, DEPTH_START, syn.PC_RANGE)
    assert tuple(xs[0].shape) == (12, 192, 116, 200) and tuple(ms[0].shape) == (1, 12, 116, 200)
    xo, mo = pe_oracle.frustum_pe_input(shapes, metas, 64, DEPTH_START, syn.PC_RANGE)
    _check(xs, ms, xo, mo)
    assert 0 < int(ms[0].sum()) < ms[0].numel()                                    # both outcomes occur
    enc = torch.nn.Sequential(torch.nn.Conv2d(192, 64, 1), torch.nn.ReLU(), torch.nn.Conv2d(64, 32, 1)).cuda()
    feats = [torch.zeros(1, 12, 4, H, W, device="cuda") for H, W in shapes]
    embs, masks = frustum.position_embeding(feats, metas, None, position_encoder=enc, depth_num=64,
                                            depth_start=DEPTH_START, pc_range=syn.PC_RANGE, embed_dims=32)
    assert tuple(embs[1].shape) == (1, 12, 32, 58, 100) and torch.equal(masks[1], ms[1])
    want = enc(xs[1]).view(1, 12, 32, 58, 100)
    assert torch.equal(embs[1], want)


def test_frustum_refuses_cpu():
    with pytest.raises(RuntimeError):
        frustum.frustum_position_input(SHAPES, syn.make_img_metas(1, 1), 8, 1, syn.PC_RANGE, device="cpu")
