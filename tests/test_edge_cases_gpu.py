"""GPU edge cases: nothing visible, single / ragged query counts, maximum sizes the ABI
accepts (L*P = 64, 24 cameras), unusual level counts, batch > 1 in wide mode."""
import numpy as np
import pytest
import torch

from graph_detr4d_b200 import ops, synthetic as syn
from graph_detr4d_b200.ops import MODE_A, MODE_C, XViewConfig
from oracle import xview_oracle as xo
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _leaf(t):
    return t.clone().requires_grad_(True)


def _run_c(sc, P, wide, Hh=8, dtype=torch.float32, seed=6):
    logits, offsets, cam = H.rand_inputs_c(sc, Hh=Hh, P=P, seed=seed)
    feats_o = [_leaf(f) for f in sc["feats"]]
    ref_o, log_o, off_o, cam_o = _leaf(sc["ref"]), _leaf(logits), _leaf(offsets), _leaf(cam)
    fn = xo.xview_c_wide_core if wide else xo.xview_c_core
    res_o = fn(feats_o, ref_o, off_o, log_o, cam_o, sc["l2i"], syn.PC_RANGE, 900, 1600, Hh)
    out_o = res_o[0]
    g = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(3))
    out_o.backward(g)
    feats_g = [_leaf(f.cuda()) for f in sc["feats"]]
    ref_g, log_g, off_g, cam_g = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
    packed = ops.pack_features(feats_g, dtype)
    cfg = XViewConfig(MODE_C, Hh, P, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=wide)
    res = ops.xview_attention(cfg, packed, ref_g, log_g, off_g, cam_g, sc["l2i"].cuda())
    out = res[0] if wide else res
    out.backward(g.cuda())
    assert H.rel_err(out.detach().cpu(), out_o.detach()) <= 1e-5
    for a, b in ((log_g, log_o), (off_g, off_o), (cam_g, cam_o), (ref_g, ref_o)):
        assert H.rel_err(a.grad.cpu(), b.grad) <= 2e-4
    for a, b in zip(feats_g, feats_o):
        assert H.rel_err(a.grad.cpu(), b.grad) <= 2e-4


@pytest.mark.parametrize("Q", [1, 7, 9])
@pytest.mark.parametrize("wide", [False, True])
def test_ragged_query_counts(Q, wide):
    _run_c(H.scene(B=1, T=1, Q=Q), P=4, wide=wide)


@pytest.mark.parametrize("wide", [False, True])
def test_maximum_points_and_cameras(wide):
    # P=16 -> L*P = 64 logits per head (the shared-memory limit), T=4 -> 24 cameras
    _run_c(H.scene(B=1, T=4, Q=12), P=16, wide=wide)


def test_too_many_points_is_refused_with_status():
    sc = H.scene(B=1, T=1, Q=4)
    logits, offsets, cam = H.rand_inputs_c(sc, P=17)
    packed = ops.pack_features([f.cuda() for f in sc["feats"]])
    cfg = XViewConfig(MODE_C, 8, 17, tuple(syn.PC_RANGE), 900.0, 1600.0)
    with pytest.raises(RuntimeError, match="status -5"):
        ops.xview_forward(cfg, packed.levels, 1, 6, sc["ref"].cuda(), logits.cuda(), offsets.cuda(),
                          cam.cuda(), sc["l2i"].cuda())


@pytest.mark.parametrize("wide", [False, True])
def test_batch_of_three_scenes(wide):
    _run_c(H.scene(B=3, T=1, Q=21), P=2, wide=wide)


@pytest.mark.parametrize("shapes", [[(9, 14)], [(13, 21), (7, 11), (4, 6), (2, 3), (1, 2), (1, 1)]])
def test_one_and_six_levels_odd_sizes(shapes):
    _run_c(H.scene(B=1, T=1, Q=30, shapes=shapes), P=3, wide=True)
    sc = H.scene(B=1, T=1, Q=30, shapes=shapes)
    logits = H.rand_inputs_a(sc)
    ref_out, ref_mask = xo.xview_a_core(sc["feats"], sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
    packed = ops.pack_features([f.cuda() for f in sc["feats"]])
    cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out, mask = ops.xview_forward(cfg, packed.levels, 1, 6, sc["ref"].cuda(), logits.cuda(),
                                  lidar2img=sc["l2i"].cuda(), want_mask=True)
    assert torch.equal(mask.cpu().bool(), ref_mask) and H.rel_err(out.cpu(), ref_out) <= 1e-5


@pytest.mark.parametrize("mode", ["A", "C", "Cwide"])
def test_nothing_visible_gives_exact_zeros(mode):
    """Every camera looks away (negative depth): mask all-false, output and every gradient 0."""
    sc = H.scene(B=1, T=1, Q=40)
    l2i = sc["l2i"].clone()
    l2i[:, :, 2, :] = torch.tensor([0.0, 0.0, 0.0, -1.0])         # cz = -1 for every point
    feats_g = [_leaf(f.cuda()) for f in sc["feats"]]
    packed = ops.pack_features(feats_g)
    ref_g = _leaf(sc["ref"].cuda())
    if mode == "A":
        logits = _leaf(H.rand_inputs_a(sc).cuda())
        cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
        out = ops.xview_attention(cfg, packed, ref_g, logits, lidar2img=l2i.cuda())
        _, mask = ops.xview_forward(cfg, packed.levels, 1, 6, ref_g.detach(), logits.detach(),
                                    lidar2img=l2i.cuda(), want_mask=True)
    else:
        lg, off, cam = (_leaf(t.cuda()) for t in H.rand_inputs_c(sc))
        logits = lg
        cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=(mode == "Cwide"))
        res = ops.xview_attention(cfg, packed, ref_g, lg, off, cam, l2i.cuda())
        out = res[0] if mode == "Cwide" else res
        _, mask = ops.xview_forward(cfg, packed.levels, 1, 6, ref_g.detach(), lg.detach(), off.detach(),
                                    cam.detach(), l2i.cuda(), want_mask=True)
    out.sum().backward()
    assert int(mask.sum()) == 0
    assert float(out.abs().max()) == 0.0
    assert float(ref_g.grad.abs().max()) == 0.0 and float(logits.grad.abs().max()) == 0.0
    assert all(float(f.grad.abs().max()) == 0.0 for f in feats_g)


def test_points_on_the_image_border_use_zero_padding():
    """Reference points projected just inside the border sample partly outside the maps:
    out-of-map corners must contribute exactly nothing (grid_sample zeros padding)."""
    sc = H.scene(B=1, T=1, Q=400, shapes=[(4, 7), (2, 3)])
    # squeeze the points towards the frustum edges: large lateral range, near depth
    ref = sc["ref"].clone()
    ref[..., 2] = 0.55
    logits = H.rand_inputs_a(sc)
    ref_out, ref_mask = xo.xview_a_core(sc["feats"], ref, logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
    packed = ops.pack_features([f.cuda() for f in sc["feats"]])
    cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out, mask = ops.xview_forward(cfg, packed.levels, 1, 6, ref.cuda(), logits.cuda(),
                                  lidar2img=sc["l2i"].cuda(), want_mask=True)
    assert torch.equal(mask.cpu().bool(), ref_mask) and ref_mask.any()
    assert H.rel_err(out.cpu(), ref_out) <= 1e-5
