"""GPU: behaviours the reference has and a fused re-implementation can silently lose (round-1 review):

  * eval-mode self-attention at B > 1 must not apply dropout (generic path of glue.self_attention)
  * two forwards before one backward (teacher/student, multi-sample losses) -- no in-place refresh of
    anything an earlier forward saved for its backward
  * a pack made under no_grad must not serve a later training forward on the same tensors
  * zeros padding never READS outside the map: non-finite pixels next to the border do not leak into
    samples whose out-of-map corners carry weight 0 (F.grid_sample / mmcv MSDA semantics)
  * variants A and V2 pass the sampled features through torch.nan_to_num BEFORE weighting
    (detr3d_transformer.py:378, :619)
  * Deform3DCrossAttn at B > 1: the reference's logits/image pairing is only self-consistent for B == 1;
    the module raises unless allow_batched=True
  * Hungarian match cost: labels outside [0, num_classes) are rejected, never read out of bounds
"""
import pytest
import torch

import graph_detr4d_b200 as g
from graph_detr4d_b200 import ops, synthetic as syn
from graph_detr4d_b200.decoder import Detr3DTransformer, Detr3DTransformerDecoder, SelfAttention
from graph_detr4d_b200.ops import MODE_A, MODE_C, MODE_V2, XViewConfig
from oracle import xview_oracle as xo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_eval_self_attention_batch2_is_deterministic_and_matches_torch():
    torch.manual_seed(3)
    sa = SelfAttention(256, 8, dropout=0.1).cuda().eval()
    q = torch.randn(50, 2, 256, device="cuda")
    pos = torch.randn(50, 2, 256, device="cuda")
    with torch.no_grad():
        y1, y2 = sa(q, pos), sa(q, pos)
        want = q + sa.attn(q + pos, q + pos, value=q, need_weights=False)[0]
    assert torch.equal(y1, y2)                                   # dropout in eval would make these differ
    assert H.rel_err(y1, want) <= 1e-5
    sa.train()                                                   # and in training it IS applied
    with torch.no_grad():
        assert not torch.equal(sa(q, pos), sa(q, pos))


def _decoder(N, layers=2, Q=40):
    torch.manual_seed(21)
    cfg = dict(type="Deform3DCrossAttn", num_cams=N, num_points=4, pc_range=syn.PC_RANGE, dropout=0.0)
    dec = Detr3DTransformerDecoder(cfg, num_layers=layers, dropout=0.0)
    model = Detr3DTransformer(dec, num_query=Q)
    for i, layer in enumerate(dec.layers):
        syn.randomize_generators(layer.attentions[1], seed=30 + i)
    return model.cuda()


def test_two_forwards_before_one_backward():
    """Different scenes (different lidar2img) through the same decoder, then ONE backward of the sum."""
    model = _decoder(6)
    sc1, sc2 = H.scene(B=1, T=1, Q=40, seed=0), H.scene(B=1, T=1, Q=40, seed=5)
    metas2 = syn.make_img_metas(2, 1)[1:]                        # shifted ego pose: other matrices
    f1 = [f.cuda().requires_grad_(True) for f in sc1["feats"]]
    f2 = [f.cuda().requires_grad_(True) for f in sc2["feats"]]
    g.clear_caches()
    s1, _, _ = model(f1, sc1["metas"], 1)
    s2, _, _ = model(f2, metas2, 1)
    w = torch.randn(s1.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    ((s1 + s2) * w).sum().backward()                             # must not raise "modified by an inplace operation"
    both = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    gf1 = [f.grad.clone() for f in f1]
    # the same two losses, one at a time
    model.zero_grad(set_to_none=True)
    for f in f1 + f2:
        f.grad = None
    g.clear_caches()
    s1, _, _ = model(f1, sc1["metas"], 1)
    (s1 * w).sum().backward()
    s2, _, _ = model(f2, metas2, 1)
    (s2 * w).sum().backward()
    single = [p.grad for p in model.parameters() if p.grad is not None]
    assert len(both) == len(single)
    for a, b in zip(both, single):
        assert H.rel_err(a, b) <= 2e-4
    for a, b in zip(gf1, [f.grad for f in f1]):
        assert H.rel_err(a, b) <= 2e-4


def test_no_grad_forward_does_not_poison_the_next_training_forward():
    sc = H.scene(B=1, T=1, Q=40)
    m = g.Deform3DCrossAttn(num_cams=6, num_points=4, pc_range=syn.PC_RANGE, dropout=0.0)
    syn.randomize_generators(m)
    m = m.cuda()
    feats = [f.cuda().requires_grad_(True) for f in sc["feats"]]
    args = dict(query_pos=sc["query_pos"].cuda(), reference_points=sc["ref"].cuda(), img_metas=sc["metas"])
    g.clear_caches()
    with torch.no_grad():
        m(sc["query"].cuda(), None, feats, **args)
    y = m(sc["query"].cuda(), None, feats, **args)               # same tensors, now with autograd
    y.sum().backward()
    assert all(f.grad is not None and float(f.grad.abs().sum()) > 0 for f in feats)


@pytest.mark.parametrize("mode,wide", [("A", False), ("C", False), ("C", True)])
def test_non_finite_pixels_outside_the_footprint_do_not_leak(mode, wide):
    """Poison the pixels on the image border ring; samples whose bilinear footprint hangs over the edge
    give those pixels weight > 0 (the reference also returns non-finite there), but a sample one pixel
    further in must stay finite, and -- the regression -- so must every sample whose out-of-map corner is
    CLAMPED onto a border pixel with weight 0.  Compare non-finiteness patterns with the oracle."""
    sc = H.scene(B=1, T=1, Q=400, seed=3)
    feats = [f.clone() for f in sc["feats"]]
    inf = float("nan") if mode == "A" else float("inf")         # A: +-FLT_MAX sums would overflow order-dependently
    for f in feats:
        f[..., 0, :] = inf
        f[..., -1, :] = float("nan")
        f[..., :, 0] = -inf
        f[..., :, -1] = inf
    packed = ops.pack_features([f.cuda() for f in feats])
    ref, l2i = sc["ref"].cuda(), sc["l2i"].cuda()
    if mode == "A":
        logits = H.rand_inputs_a(sc)
        want, _ = xo.xview_a_core(feats, sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
        cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
        out, _ = ops.xview_forward(cfg, packed.levels, 1, 6, ref, logits.cuda(), lidar2img=l2i)
        # variant A: nan_to_num -> NaN samples count as 0, Inf as +-FLT_MAX (test_variant_a_nan_to_num_on_samples)
    else:
        logits, offsets, cam = H.rand_inputs_c(sc)
        cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=wide)
        res, _ = ops.xview_forward(cfg, packed.levels, 1, 6, ref, logits.cuda(), offsets.cuda(), cam.cuda(), l2i)
        if wide:
            out = res[0]
            want, _ = xo.xview_c_wide_core(feats, sc["ref"], offsets, logits, cam, sc["l2i"], syn.PC_RANGE,
                                           900, 1600, 8)
        else:
            out = res
            want, _ = xo.xview_c_core(feats, sc["ref"], offsets, logits, cam, sc["l2i"], syn.PC_RANGE, 900, 1600, 8)
    out = out.cpu()
    fin_o, fin_k = torch.isfinite(want), torch.isfinite(out)
    # never MORE non-finite outputs than the reference semantics produce
    assert not (fin_o & ~fin_k).any(), int((fin_o & ~fin_k).sum())
    # (the reference ALSO turns every masked-out point that touches a poisoned pixel into NaN -- 0 * Inf:
    # it samples first and masks after -- so most of its outputs are NaN here; the kernel never samples
    # masked-out points.  Compare where both are finite.)
    ok = fin_o & fin_k & (want.abs() < 1e30)
    assert int(ok.sum()) > 2000
    assert H.rel_err(out[ok], want[ok]) <= 1e-5


def test_variant_a_nan_to_num_on_samples():
    """detr3d_transformer.py:378: a NaN sample contributes 0, not NaN; the mask and the other samples
    of the same query are unaffected."""
    sc = H.scene(B=1, T=1, Q=300, seed=4)
    feats = [f.clone() for f in sc["feats"]]
    feats[1][0, 2, 7] = float("nan")                            # camera 2, level 1, channel 7: all pixels
    feats[3][0, 0, 100] = float("inf")                           # camera 0, level 3, channel 100
    logits = H.rand_inputs_a(sc)
    want, _ = xo.xview_a_core(feats, sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
    packed = ops.pack_features([f.cuda() for f in feats])
    cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out, _ = ops.xview_forward(cfg, packed.levels, 1, 6, sc["ref"].cuda(), logits.cuda(), lidar2img=sc["l2i"].cuda())
    out = out.cpu()
    assert torch.isfinite(want[..., 7]).all() and torch.isfinite(out[..., 7]).all()
    assert H.rel_err(out[..., 7], want[..., 7]) <= 1e-5
    # +-Inf samples become +-FLT_MAX times a sigmoid weight: compare where the reference stays finite
    both = torch.isfinite(want) & torch.isfinite(out)
    assert torch.equal(torch.isfinite(want), torch.isfinite(out))
    big = want.abs() > 1e30
    assert big.any()
    assert H.rel_err(out[both & ~big], want[both & ~big]) <= 1e-5
    assert float(((out[both & big] - want[both & big]).abs() / want[both & big].abs()).max()) <= 1e-5


def test_deform3d_module_raises_for_batches_unless_allowed():
    sc = H.scene(B=2, T=1, Q=30)
    kw = dict(num_cams=6, num_points=4, pc_range=syn.PC_RANGE, dropout=0.0)
    torch.manual_seed(2)
    m = g.Deform3DCrossAttn(**kw)
    syn.randomize_generators(m)
    m = m.cuda().eval()
    args = (sc["query"].cuda(), None, [f.cuda() for f in sc["feats"]])
    kwargs = dict(query_pos=sc["query_pos"].cuda(), reference_points=sc["ref"].cuda(), img_metas=sc["metas"])
    g.clear_caches()
    with pytest.raises(ValueError, match="allow_batched"):
        m(*args, **kwargs)
    m.allow_batched = True
    g.clear_caches()
    with torch.no_grad():
        y = m(*args, **kwargs)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    want = xo.deform3d_cross_attn_forward(sd, sc["query"], sc["feats"], sc["query_pos"], sc["ref"], sc["metas"],
                                          syn.PC_RANGE, 8, reference_batch_order=False)   # per-sample logits
    assert H.rel_err(y.cpu(), want) <= 5e-5
    quirk = xo.deform3d_cross_attn_forward(sd, sc["query"], sc["feats"], sc["query_pos"], sc["ref"], sc["metas"],
                                           syn.PC_RANGE, 8, reference_batch_order=True)   # what the reference does
    assert H.rel_err(y.cpu(), quirk) > 1e-3                      # documented: NOT the reference's B > 1 pairing


def test_assigner_rejects_out_of_range_labels():
    from graph_detr4d_b200.assign import BatchedHungarianAssigner3D
    gen = torch.Generator().manual_seed(0)
    bp = torch.randn(2, 1, 50, 10, generator=gen).cuda()
    cs = torch.randn(2, 1, 50, 10, generator=gen).cuda()
    gt = torch.rand(4, 9, generator=gen).cuda() + 0.5
    a = BatchedHungarianAssigner3D()
    for bad in (-1, 10):
        lab = torch.tensor([1, bad, 3, 0]).cuda()
        with pytest.raises(ValueError, match="label"):
            a.assign_layers(bp, cs, [gt], [lab])
    with pytest.raises(ValueError, match="labels for"):
        a.assign_layers(bp, cs, [gt], [torch.tensor([1, 2, 3]).cuda()])
    inds, labels = a.assign_layers(bp, cs, [gt], [torch.tensor([1, 2, 3, 0]).cuda()])
    assert int((inds > 0).sum()) == 2 * 4
