"""GPU, row f4 (second half): the fused position-embedding block (include/gd4d_fpe.h,
graph_detr4d_b200/fpe.py) vs the CPU oracle and vs the golden fixture frozen from the reference's own
lines (detr3d_head_pe.py:510-553).  Masks bit-exact; values 1e-5 of max|ref| (CUDA sinf/cosf/expf vs the
host's libm, cuDNN/cuBLAS 1x1 convolutions vs the host's); the 3 masked bottom rows of the real
928x1600 geometry -- whose sine argument is ~ -3e6 -- are checked separately, bit-identical arguments."""
import types

import pytest
import torch
import torch.nn as nn

from graph_detr4d_b200 import fpe, synthetic as syn
from oracle import fpe_oracle
from tests import helpers as H
from tests.test_fpe_oracle import load_fpe_golden

pytestmark = pytest.mark.gpu


class _SE(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_reduce = nn.Conv2d(c, c, 1, bias=True)
        self.act1 = nn.ReLU()
        self.conv_expand = nn.Conv2d(c, c, 1, bias=True)
        self.gate = nn.Sigmoid()


def _head(C, D, F_, sd=None, with_detach=True):
    h = nn.Module()
    h.position_encoder = nn.Sequential(nn.Conv2d(3 * D, 4 * C, 1), nn.ReLU(), nn.Conv2d(4 * C, C, 1))
    h.adapt_pos3d = nn.Sequential(nn.Conv2d(C * 3 // 2, 4 * C, 1), nn.ReLU(), nn.Conv2d(4 * C, C, 1))
    h.fpe = _SE(C)
    if sd is not None:
        h.load_state_dict(sd, strict=True)                      # same parameter names as the reference head
    h.positional_encoding = types.SimpleNamespace(num_feats=F_, temperature=10000, normalize=True,
                                                  scale=2 * 3.141592653589793, eps=1e-6, offset=-0.5)
    h.depth_num, h.depth_start, h.pc_range, h.with_detach = D, 1, syn.PC_RANGE, with_detach
    return h.cuda()


def test_block_matches_reference_golden():
    torch.backends.cudnn.allow_tf32 = False
    gd = load_fpe_golden()
    head = _head(gd["C"], gd["D"], gd["F"], gd["sd"])
    feats = [f.cuda().requires_grad_(True) for f in gd["feats"]]
    outs = fpe.position_embed_features(head, feats, gd["metas"])
    sum((o * g.cuda()).sum() for o, g in zip(outs, gd["gouts"])).backward()
    masks = fpe.level_masks(gd["shapes"], gd["metas"], 6 * gd["T"])
    for l in range(len(outs)):
        assert torch.equal(masks[l].cpu(), gd["masks"][l])
        assert H.rel_err(outs[l].detach().cpu(), gd["outs"][l]) <= 1e-5
        assert H.rel_err(feats[l].grad.cpu(), gd["grads"][l]) <= 1e-5
    assert float(feats[0].grad[:, 6:].abs().max()) == 0.0 and float(feats[1].grad[:, 6:].abs().max()) > 0.0


@pytest.mark.parametrize("B,T", [(1, 2), (2, 1)])
def test_masks_and_sine_embedding_full_geometry(B, T):
    """928x1600 padded / 900x1600 images, the real level sizes: masks bit-exact; sine embedding vs the
    oracle -- tight where the argument is O(1), and sin/cos-accurate (same fp32 argument) on masked rows."""
    metas = syn.make_img_metas(B, T)
    N = 6 * T
    shapes = H.FULL_SHAPES
    want_masks = fpe_oracle.level_masks(B, N, shapes, metas)
    got_masks = fpe.level_masks(shapes, metas, N)
    for l, (h, w) in enumerate(shapes):
        assert torch.equal(got_masks[l].cpu(), want_masks[l])
        assert bool(want_masks[l].any()) == (l < 2)             # 900 of 928 rows: only strides 8 and 16 see padding rows
        got = fpe.sine_pe3d((h, w), metas, N, 128, offset=-0.5).cpu().view(B, N, 384, h, w)
        want = fpe_oracle.sine_pe3d(want_masks[l], 128, offset=-0.5)
        assert float((got - want).abs().max()) <= 2e-6          # |sin|,|cos| <= 1: absolute


def test_sine_embedding_ragged_image_sizes_and_combine_backward():
    gd = load_fpe_golden()
    N = 6 * gd["T"]
    for l, (h, w) in enumerate(gd["shapes"]):
        got = fpe.sine_pe3d((h, w), gd["metas"], N, gd["F"], offset=-0.5).cpu().view(1, N, 3 * gd["F"], h, w)
        want = fpe_oracle.sine_pe3d(gd["masks"][l], gd["F"], offset=-0.5)
        assert float((got - want).abs().max()) <= 2e-6
    g = torch.Generator().manual_seed(3)
    a, b, c, d = (torch.randn(2, 5, 7, 9, generator=g).cuda().requires_grad_(True) for _ in range(4))
    y = fpe.fpe_combine(a, b, c, d)
    w = torch.randn(y.shape, generator=g).cuda()
    (y * w).sum().backward()
    a2, b2, c2, d2 = (t.detach().clone().requires_grad_(True) for t in (a, b, c, d))
    y2 = a2 + (b2 * c2.sigmoid() + d2)
    (y2 * w).sum().backward()
    assert H.rel_err(y, y2) <= 1e-6
    for p, q in ((a, a2), (b, b2), (c, c2), (d, d2)):
        assert H.rel_err(p.grad, q.grad) <= 1e-6


def test_cpu_tensors_are_refused():
    gd = load_fpe_golden()
    head = _head(gd["C"], gd["D"], gd["F"], gd["sd"]).cpu()
    with pytest.raises(RuntimeError):
        fpe.position_embed_features(head, gd["feats"], gd["metas"])
