"""Parity at BASELINE.json's FULL sizes (900 queries, 4 levels of 256 channels at 928x1600,
6 / 12 cameras, fp32 and bf16 maps) through size-independent properties -- the CPU oracle
needs minutes there, these need none:

  * linearity in the feature maps            out(a f1 + b f2) = a out(f1) + b out(f2)
  * adjointness (dot-product test)           <out(f), g> = <f, grad_f(g)>   (forward vs backward kernel)
  * all-ones maps                            out = wsum in every channel (bilinear weights sum to the
                                             in-bounds weight; ties the value path to the bias path)
  * camera permutation                       permuting cameras, matrices and camera logits together
                                             leaves the output unchanged up to summation order
  * softmax shift invariance                 sum_j dL/dlogit_j = 0 per (query, head)
  * chain rule of X = ref*span + lo + off    dL/dref = span * sum_{h,p} dL/doffset
  * valid-projection count                   the kernel's mask agrees with the roofline bookkeeping
  * bit-reproducible forward
"""
import pytest
import torch

from graph_detr4d_b200 import ops, roofline, synthetic as syn
from graph_detr4d_b200.ops import MODE_A, MODE_C, XViewConfig
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _scene(T, dtype, seed=0):
    sc = H.scene(B=1, T=T, Q=900, shapes=H.FULL_SHAPES, seed=seed)
    feats = [f.cuda() for f in sc["feats"]]
    levels = ops.pack_features(feats, dtype).levels
    logits, offsets, cam = (t.cuda() for t in H.rand_inputs_c(sc, off_std=1.5))
    return sc, levels, logits, offsets, cam


@pytest.mark.parametrize("T,dtype,wide", [(1, torch.float32, True), (2, torch.float32, True),
                                          (2, torch.bfloat16, True), (2, torch.float32, False)])
def test_mode_c_full_size_properties(T, dtype, wide):
    sc, levels, logits, offsets, cam = _scene(T, dtype)
    N, Q, Hh, P = sc["N"], 900, 8, 4
    ref, l2i = sc["ref"].cuda(), sc["l2i"].cuda()
    cfg = XViewConfig(MODE_C, Hh, P, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=wide)
    args = (ref, logits, offsets, cam, l2i)

    def fwd(vals):
        res, _ = ops.xview_forward(cfg, vals, 1, N, *args)
        return res if wide else (res, None)

    out1, ws1 = fwd(levels)
    out1b, _ = fwd(levels)
    assert torch.equal(out1, out1b)                                        # bit-reproducible

    # ---- linearity in the maps (fp32 maps only: a*f1+b*f2 is not representable in bf16) ----
    if dtype == torch.float32:
        g = torch.Generator(device="cuda").manual_seed(3)
        other = [torch.randn(v.shape, device="cuda", generator=g) for v in levels]
        out2, _ = fwd(other)
        mix = [0.75 * a - 1.5 * b for a, b in zip(levels, other)]
        out3, _ = fwd(mix)
        assert _rel(out3, 0.75 * out1 - 1.5 * out2) <= 1e-5

    # ---- all-ones maps: every channel of the aggregate equals wsum --------------------------
    ones = [torch.ones_like(v) for v in levels]
    o1, w1 = fwd(ones)
    if wide:
        assert torch.equal(w1, ws1)                                        # wsum does not depend on the maps
        assert _rel(o1, w1.unsqueeze(-1).expand_as(o1)) <= 1e-5
    else:
        oh = o1.view(1, Q, Hh, -1)                                         # every channel of a head slice
        assert float((oh - oh[..., :1]).abs().max()) <= 1e-5 * float(o1.abs().max())
    # ---- adjointness: <out(f), g> == <f, grad_f(g)> -------------------------------------------
    gen = torch.Generator(device="cuda").manual_seed(9)
    gout = torch.randn(out1.shape, device="cuda", generator=gen)
    gws = torch.zeros_like(ws1) if wide else None
    gv = [torch.zeros(v.shape, device="cuda", dtype=torch.float32) for v in levels]
    g_attn, g_off, g_cam, g_ref = ops.xview_backward(cfg, levels, 1, N, *args, gout, gv, grad_wsum=gws)
    lhs = float((out1.double() * gout.double()).sum())
    rhs = float(sum((v.double() * gg.double()).sum() for v, gg in zip(levels, gv)))
    scale = float((out1.double() * gout.double()).abs().sum())
    assert abs(lhs - rhs) <= 2e-5 * scale

    # ---- softmax shift invariance and the ref/offset chain rule --------------------------------
    per_head = g_attn.view(1, Q, Hh, -1).sum(-1)
    assert float(per_head.abs().max()) <= 2e-4 * float(g_attn.abs().max()) * g_attn.shape[-1] ** 0.5
    span = torch.tensor([syn.PC_RANGE[3] - syn.PC_RANGE[0], syn.PC_RANGE[4] - syn.PC_RANGE[1],
                         syn.PC_RANGE[5] - syn.PC_RANGE[2]], device="cuda")
    want_ref = g_off.view(1, Q, Hh * P, 3).sum(2) * span
    assert _rel(g_ref, want_ref) <= 2e-4

    # ---- the mask agrees with the host-side bookkeeping used for the roofline -------------------
    _, mask = ops.xview_forward(cfg, levels, 1, N, *args, want_mask=True)
    stats = roofline.count_corner_reads(MODE_C, [(int(v.shape[1]), int(v.shape[2])) for v in levels], ref,
                                        offsets, l2i, syn.PC_RANGE, 900.0, 1600.0, Hh, P)
    assert abs(int(mask.sum()) - stats["valid_samples"]) <= max(2, int(1e-5 * stats["valid_samples"]))
    assert 0.10 < stats["valid_fraction"] < 0.30

    # ---- camera permutation ------------------------------------------------------------------------
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(4))
    lv_p = []
    for v in levels:
        lv_p.append(v.view(1, N, *v.shape[1:])[:, perm.cuda()].reshape(v.shape).contiguous())
    l2i_p = l2i[:, perm.cuda()].contiguous()
    # camera logits: the kernel reads weight(n,q) = flat[n*Q+q] of the (B,Q,N) tensor (reference view quirk)
    cam_p = cam.reshape(1, N, Q)[:, perm.cuda()].reshape(1, Q, N).contiguous()
    res_p, _ = ops.xview_forward(cfg, lv_p, 1, N, ref, logits, offsets, cam_p, l2i_p)
    out_p = res_p[0] if wide else res_p
    assert _rel(out_p, out1) <= 1e-5


def test_mode_a_full_size_properties():
    sc = H.scene(B=1, T=1, Q=900, shapes=H.FULL_SHAPES)
    levels = ops.pack_features([f.cuda() for f in sc["feats"]]).levels
    logits = H.rand_inputs_a(sc).cuda()
    ref, l2i = sc["ref"].cuda(), sc["l2i"].cuda()
    cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out1, mask = ops.xview_forward(cfg, levels, 1, 6, ref, logits, lidar2img=l2i, want_mask=True)
    g = torch.Generator(device="cuda").manual_seed(3)
    other = [torch.randn(v.shape, device="cuda", generator=g) for v in levels]
    out2, _ = ops.xview_forward(cfg, other, 1, 6, ref, logits, lidar2img=l2i)
    out3, _ = ops.xview_forward(cfg, [2.0 * a + 0.5 * b for a, b in zip(levels, other)], 1, 6, ref, logits,
                                lidar2img=l2i)
    assert _rel(out3, 2.0 * out1 + 0.5 * out2) <= 1e-5
    # queries no camera sees produce exact zeros
    unseen = mask.sum(-1) == 0
    assert int(unseen.sum()) > 0 and float(out1[unseen].abs().max()) == 0.0
    # adjointness
    gout = torch.randn(out1.shape, device="cuda", generator=g)
    gv = [torch.zeros_like(v) for v in levels]
    ops.xview_backward(cfg, levels, 1, 6, ref, logits, None, None, l2i, gout, gv)
    lhs = float((out1.double() * gout.double()).sum())
    rhs = float(sum((v.double() * gg.double()).sum() for v, gg in zip(levels, gv)))
    assert abs(lhs - rhs) <= 2e-5 * float((out1.double() * gout.double()).abs().sum())
