"""GPU parity AGAINST THE CPU ORACLE at BASELINE.json's full shapes: 900 queries, 4 FPN levels of 256
channels at 928x1600 (116x200 ... 15x25), 6 and 12 cameras -- the shapes bench.py times.

The oracle (oracle/xview_oracle.py, pinned to the executed reference by tests/test_oracle_vs_reference.py)
needs 3-6 s per case on the host here, so these are plain value comparisons, not property tests:

  * wide fp32  N=6   the bench workload's kernel (configs[1])
  * wide fp32  N=12  Graph-DETR4D T=2 (configs[2] shapes, fp32 maps)
  * wide bf16  N=12  configs[2]: vs the oracle fed the SAME bf16-rounded maps, and vs the fp32 oracle
  * narrow fp32 N=12 the mmcv op boundary (projected value, head slices)
  * mode A fp32 N=6  configs[0]
  * the Deform3DCrossAttn MODULE (packed generator GEMM -> gen_stride kernels -> per-head W_v GEMM ->
    output_proj -> position encoder) against the oracle's whole-module restatement

Tolerances (north_star): forward <= 1e-5 * max|ref| (fp32, and bf16 vs same-rounded maps), 8e-3 for bf16
maps vs the fp32 oracle, projection mask bit-exact, every gradient <= 2e-4 * max|ref| (atomic order).
The module adds six cuBLAS fp32 GEMMs whose summation order differs from the host GEMMs: 2e-5, stated.
"""
import pytest
import torch

import graph_detr4d_b200 as g
from graph_detr4d_b200 import ops, synthetic as syn
from graph_detr4d_b200.ops import MODE_A, MODE_C, XViewConfig
from oracle import xview_oracle as xo
from tests import helpers as H

pytestmark = pytest.mark.gpu
FWD_TOL, BF16_TOL, GRAD_TOL = 1e-5, 8e-3, 2e-4
Q, HH, P = 900, 8, 4


def _leaf(t):
    return t.clone().requires_grad_(True)


@pytest.mark.parametrize("T,dtype", [(1, torch.float32), (2, torch.float32), (2, torch.bfloat16)])
def test_wide_kernels_vs_oracle_full_shapes(T, dtype):
    sc = H.scene(B=1, T=T, Q=Q, shapes=H.FULL_SHAPES)
    N = sc["N"]
    logits, offsets, cam = H.rand_inputs_c(sc, P=P, off_std=1.5)
    gen = torch.Generator().manual_seed(21)
    g1 = torch.randn(1, HH, Q, 256, generator=gen)
    g2 = torch.randn(1, HH, Q, generator=gen)
    feats_src = [f.to(dtype).float() for f in sc["feats"]]                  # bf16: the same rounded maps
    feats_o = [_leaf(f) for f in feats_src]
    ref_o, log_o, off_o, cam_o = _leaf(sc["ref"]), _leaf(logits), _leaf(offsets), _leaf(cam)
    agg_o, ws_o, mask_o = xo.xview_c_wide_core(feats_o, ref_o, off_o, log_o, cam_o, sc["l2i"], syn.PC_RANGE,
                                               900, 1600, HH, return_mask=True)
    ((agg_o * g1).sum() + (ws_o * g2).sum()).backward()

    feats_g = [_leaf(f.cuda()) for f in feats_src]
    ref_g, log_g, off_g, cam_g = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
    packed = ops.pack_features(feats_g, dtype)
    assert packed.levels[0].dtype == dtype and packed.shapes == [tuple(s) for s in H.FULL_SHAPES]
    cfg = XViewConfig(MODE_C, HH, P, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    l2i = sc["l2i"].cuda()
    agg, ws = ops.xview_attention(cfg, packed, ref_g, log_g, off_g, cam_g, l2i)
    ((agg * g1.cuda()).sum() + (ws * g2.cuda()).sum()).backward()
    _, mask = ops.xview_forward(cfg, packed.levels, 1, N, sc["ref"].cuda(), logits.cuda(), offsets.cuda(),
                                cam.cuda(), l2i, want_mask=True)
    assert torch.equal(mask.cpu().bool(), mask_o[:, :, :, :, 0, :]), "projection mask must be bit-exact"
    assert 0.10 < float(mask.float().mean()) < 0.30
    assert H.rel_err(agg.detach().cpu(), agg_o.detach()) <= FWD_TOL
    assert H.rel_err(ws.detach().cpu(), ws_o.detach()) <= FWD_TOL
    assert H.rel_err(log_g.grad.cpu(), log_o.grad) <= GRAD_TOL
    assert H.rel_err(cam_g.grad.cpu(), cam_o.grad) <= GRAD_TOL
    assert H.rel_err(off_g.grad.cpu(), off_o.grad) <= GRAD_TOL
    assert H.rel_err(ref_g.grad.cpu(), ref_o.grad) <= GRAD_TOL
    for fg, fo in zip(feats_g, feats_o):
        assert fg.grad.dtype == torch.float32 and fg.grad.shape == fo.grad.shape
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL
    if dtype == torch.bfloat16:                                             # vs the un-rounded fp32 maps
        with torch.no_grad():
            agg_f, _ = xo.xview_c_wide_core(sc["feats"], sc["ref"], offsets, logits, cam, sc["l2i"],
                                            syn.PC_RANGE, 900, 1600, HH)
        assert H.rel_err(agg.detach().cpu(), agg_f) <= BF16_TOL


def test_narrow_kernels_vs_oracle_full_shapes():
    sc = H.scene(B=1, T=2, Q=Q, shapes=H.FULL_SHAPES)
    logits, offsets, cam = H.rand_inputs_c(sc, P=P, off_std=1.5)
    gout = torch.randn(1, Q, 256, generator=torch.Generator().manual_seed(22))
    feats_o = [_leaf(f) for f in sc["feats"]]
    ref_o, log_o, off_o, cam_o = _leaf(sc["ref"]), _leaf(logits), _leaf(offsets), _leaf(cam)
    out_o, mask_o = xo.xview_c_core(feats_o, ref_o, off_o, log_o, cam_o, sc["l2i"], syn.PC_RANGE, 900, 1600, HH)
    out_o.backward(gout)
    feats_g = [_leaf(f.cuda()) for f in sc["feats"]]
    ref_g, log_g, off_g, cam_g = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
    packed = ops.pack_features(feats_g)
    cfg = XViewConfig(MODE_C, HH, P, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out = ops.xview_attention(cfg, packed, ref_g, log_g, off_g, cam_g, sc["l2i"].cuda())
    out.backward(gout.cuda())
    assert H.rel_err(out.detach().cpu(), out_o.detach()) <= FWD_TOL
    for a, b in ((log_g, log_o), (cam_g, cam_o), (off_g, off_o), (ref_g, ref_o)):
        assert H.rel_err(a.grad.cpu(), b.grad) <= GRAD_TOL
    for fg, fo in zip(feats_g, feats_o):
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL


def test_mode_a_vs_oracle_full_shapes():
    sc = H.scene(B=1, T=1, Q=Q, shapes=H.FULL_SHAPES)
    logits = H.rand_inputs_a(sc)
    gout = torch.randn(1, Q, 256, generator=torch.Generator().manual_seed(23))
    feats_o = [_leaf(f) for f in sc["feats"]]
    ref_o, log_o = _leaf(sc["ref"]), _leaf(logits)
    out_o, mask_o = xo.xview_a_core(feats_o, ref_o, log_o, sc["l2i"], syn.PC_RANGE, 900, 1600)
    out_o.backward(gout)
    feats_g = [_leaf(f.cuda()) for f in sc["feats"]]
    ref_g, log_g = _leaf(sc["ref"].cuda()), _leaf(logits.cuda())
    packed = ops.pack_features(feats_g)
    cfg = XViewConfig(MODE_A, HH, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out = ops.xview_attention(cfg, packed, ref_g, log_g, lidar2img=sc["l2i"].cuda())
    out.backward(gout.cuda())
    _, mask = ops.xview_forward(cfg, packed.levels, 1, 6, sc["ref"].cuda(), logits.cuda(),
                                lidar2img=sc["l2i"].cuda(), want_mask=True)
    assert torch.equal(mask.cpu().bool(), mask_o), "projection mask must be bit-exact"
    assert H.rel_err(out.detach().cpu(), out_o.detach()) <= FWD_TOL
    assert H.rel_err(log_g.grad.cpu(), log_o.grad) <= GRAD_TOL
    assert H.rel_err(ref_g.grad.cpu(), ref_o.grad) <= GRAD_TOL
    for fg, fo in zip(feats_g, feats_o):
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL


@pytest.mark.parametrize("T,feature_dtype", [(1, None), (2, None), (2, "bf16")])
def test_module_vs_oracle_full_shapes(T, feature_dtype):
    """Whole Deform3DCrossAttn module on the bench shapes (wide kernels through the packed generator GEMM)."""
    MODULE_TOL = 2e-5      # + six cuBLAS fp32 GEMMs (K = 256) vs the host's: summation order
    sc = H.scene(B=1, T=T, Q=Q, shapes=H.FULL_SHAPES)
    torch.manual_seed(7)
    m = g.Deform3DCrossAttn(num_cams=sc["N"], num_points=P, pc_range=syn.PC_RANGE, dropout=0.1,
                            feature_dtype=feature_dtype)
    syn.randomize_generators(m)
    m = m.cuda().eval()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    feats_src = [f.to(torch.bfloat16).float() for f in sc["feats"]] if feature_dtype else sc["feats"]
    q_o, rp_o = _leaf(sc["query"]), _leaf(sc["ref"])
    y_o = xo.deform3d_cross_attn_forward(sd, q_o, feats_src, sc["query_pos"], rp_o, sc["metas"], syn.PC_RANGE, HH)
    gout = torch.randn(y_o.shape, generator=torch.Generator().manual_seed(24))
    (y_o * gout).sum().backward()
    q_g, rp_g = _leaf(sc["query"].cuda()), _leaf(sc["ref"].cuda())
    g.clear_caches()
    y = m(q_g, None, [f.cuda() for f in sc["feats"]], query_pos=sc["query_pos"].cuda(), reference_points=rp_g,
          img_metas=sc["metas"])
    (y * gout.cuda()).sum().backward()
    assert m._use_wide(g.modules._PACK_CACHE._packed)                      # the kernels bench.py times
    assert H.rel_err(y.detach().cpu(), y_o.detach()) <= MODULE_TOL
    assert H.rel_err(q_g.grad.cpu(), q_o.grad) <= GRAD_TOL
    assert H.rel_err(rp_g.grad.cpu(), rp_o.grad) <= GRAD_TOL
