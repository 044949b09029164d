"""GPU parity: the fused sm_100a kernels (through the C ABI) vs the CPU oracle.

Tolerances (BASELINE.json north_star):
  fp32 features : max|out - ref| <= 1e-5 * max|ref|
  bf16 features : same 1e-5 bound vs the oracle fed the SAME bf16-rounded features
                  (all kernel math is fp32), and <= 8e-3 * max|ref| vs the fp32 oracle
  mask          : bit-exact
  gradients     : <= 2e-4 * max|ref| (fp32 atomics reorder sums; the oracle's own
                  autograd accumulates in a different order)
"""
import pytest
import torch

from graph_detr4d_b200 import ops
from graph_detr4d_b200.ops import MODE_A, MODE_C, XViewConfig
from graph_detr4d_b200 import synthetic as syn
from oracle import xview_oracle as xo
from tests import helpers as H

pytestmark = pytest.mark.gpu
FWD_TOL = 1e-5
BF16_TOL = 8e-3
GRAD_TOL = 2e-4


def _pack(feats, dtype=None):
    return ops.pack_features([f.cuda() for f in feats], dtype)


@pytest.mark.parametrize("B,T,Q,P", [(1, 1, 128, 1), (2, 2, 77, 1), (1, 1, 64, 3)])
def test_mode_a_forward_and_mask(B, T, Q, P):
    sc = H.scene(B=B, T=T, Q=Q)
    logits = H.rand_inputs_a(sc, P=P)
    ref_out, ref_mask = xo.xview_a_core(sc["feats"], sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
    packed = _pack(sc["feats"])
    cfg = XViewConfig(MODE_A, 8, P, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out, mask = ops.xview_forward(cfg, packed.levels, B, sc["N"], sc["ref"].cuda(), logits.cuda(),
                                  lidar2img=sc["l2i"].cuda(), want_mask=True)
    assert torch.equal(mask.cpu().bool(), ref_mask), "projection mask must be bit-exact"
    assert ref_mask.any()
    assert H.rel_err(out.cpu(), ref_out) <= FWD_TOL


@pytest.mark.parametrize("B,T,Q,P", [(1, 2, 128, 4), (2, 1, 50, 4), (1, 1, 33, 1), (1, 2, 40, 8)])
def test_mode_c_forward_and_mask(B, T, Q, P):
    sc = H.scene(B=B, T=T, Q=Q)
    logits, offsets, cam = H.rand_inputs_c(sc, P=P)
    ref_out, ref_mask = xo.xview_c_core(sc["feats"], sc["ref"], offsets, logits, cam, sc["l2i"],
                                        syn.PC_RANGE, 900, 1600, 8)
    packed = _pack(sc["feats"])
    cfg = XViewConfig(MODE_C, 8, P, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out, mask = ops.xview_forward(cfg, packed.levels, B, sc["N"], sc["ref"].cuda(), logits.cuda(),
                                  offsets.cuda(), cam.cuda(), sc["l2i"].cuda(), want_mask=True)
    # oracle mask is (B,N,Q,Hh,L,P), identical over L (offsets are shared across levels)
    assert torch.equal(ref_mask[:, :, :, :, 0, :], ref_mask[:, :, :, :, -1, :])
    assert torch.equal(mask.cpu().bool(), ref_mask[:, :, :, :, 0, :]), "projection mask must be bit-exact"
    assert ref_mask.any()
    assert H.rel_err(out.cpu(), ref_out) <= FWD_TOL


@pytest.mark.parametrize("mode", ["A", "C"])
def test_bf16_features(mode):
    sc = H.scene(B=1, T=2, Q=96)
    feats_bf = [f.to(torch.bfloat16) for f in sc["feats"]]
    feats_rounded = [f.float() for f in feats_bf]
    packed = _pack(sc["feats"], torch.bfloat16)
    assert packed.levels[0].dtype == torch.bfloat16
    if mode == "A":
        logits = H.rand_inputs_a(sc)
        ref_same, _ = xo.xview_a_core(feats_rounded, sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
        ref_full, _ = xo.xview_a_core(sc["feats"], sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
        cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
        out, _ = ops.xview_forward(cfg, packed.levels, 1, sc["N"], sc["ref"].cuda(), logits.cuda(),
                                   lidar2img=sc["l2i"].cuda())
    else:
        logits, offsets, cam = H.rand_inputs_c(sc)
        ref_same, _ = xo.xview_c_core(feats_rounded, sc["ref"], offsets, logits, cam, sc["l2i"],
                                      syn.PC_RANGE, 900, 1600, 8)
        ref_full, _ = xo.xview_c_core(sc["feats"], sc["ref"], offsets, logits, cam, sc["l2i"],
                                      syn.PC_RANGE, 900, 1600, 8)
        cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0)
        out, _ = ops.xview_forward(cfg, packed.levels, 1, sc["N"], sc["ref"].cuda(), logits.cuda(),
                                   offsets.cuda(), cam.cuda(), sc["l2i"].cuda())
    assert H.rel_err(out.cpu(), ref_same) <= FWD_TOL
    assert H.rel_err(out.cpu(), ref_full) <= BF16_TOL


def _leaf(t):
    return t.clone().requires_grad_(True)


@pytest.mark.parametrize("B,T,Q,P", [(1, 1, 96, 1), (2, 1, 40, 2), (1, 1, 24, 40)])
def test_mode_a_backward(B, T, Q, P):
    sc = H.scene(B=B, T=T, Q=Q)
    logits = H.rand_inputs_a(sc, P=P)
    g = torch.Generator().manual_seed(9)
    gout = torch.randn(B, Q, sc["C"], generator=g)
    # oracle autograd
    feats_o = [_leaf(f) for f in sc["feats"]]
    ref_o, log_o = _leaf(sc["ref"]), _leaf(logits)
    out_o, _ = xo.xview_a_core(feats_o, ref_o, log_o, sc["l2i"], syn.PC_RANGE, 900, 1600)
    out_o.backward(gout)
    # product
    feats_g = [_leaf(f.cuda()) for f in sc["feats"]]
    ref_g, log_g = _leaf(sc["ref"].cuda()), _leaf(logits.cuda())
    packed = ops.pack_features(feats_g)
    cfg = XViewConfig(MODE_A, 8, P, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out = ops.xview_attention(cfg, packed, ref_g, log_g, lidar2img=sc["l2i"].cuda())
    out.backward(gout.cuda())
    assert H.rel_err(out.detach().cpu(), out_o.detach()) <= FWD_TOL
    assert H.rel_err(log_g.grad.cpu(), log_o.grad) <= GRAD_TOL
    assert H.rel_err(ref_g.grad.cpu(), ref_o.grad) <= GRAD_TOL
    for fg, fo in zip(feats_g, feats_o):
        assert fg.grad.shape == fo.grad.shape
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL


@pytest.mark.parametrize("B,T,Q,P,dtype", [(1, 2, 96, 4, torch.float32), (2, 1, 30, 2, torch.float32),
                                           (1, 2, 64, 4, torch.bfloat16)])
def test_mode_c_backward(B, T, Q, P, dtype):
    sc = H.scene(B=B, T=T, Q=Q)
    logits, offsets, cam = H.rand_inputs_c(sc, P=P)
    g = torch.Generator().manual_seed(9)
    gout = torch.randn(B, Q, sc["C"], generator=g)
    feats_src = [f.to(dtype).float() for f in sc["feats"]]      # bf16 case: same rounded features
    feats_o = [_leaf(f) for f in feats_src]
    ref_o, log_o, off_o, cam_o = _leaf(sc["ref"]), _leaf(logits), _leaf(offsets), _leaf(cam)
    out_o, _ = xo.xview_c_core(feats_o, ref_o, off_o, log_o, cam_o, sc["l2i"], syn.PC_RANGE, 900, 1600, 8)
    out_o.backward(gout)

    feats_g = [_leaf(f.cuda()) for f in feats_src]
    ref_g, log_g, off_g, cam_g = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
    packed = ops.pack_features(feats_g, dtype)
    cfg = XViewConfig(MODE_C, 8, P, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out = ops.xview_attention(cfg, packed, ref_g, log_g, off_g, cam_g, sc["l2i"].cuda())
    out.backward(gout.cuda())
    assert H.rel_err(out.detach().cpu(), out_o.detach()) <= FWD_TOL
    assert H.rel_err(log_g.grad.cpu(), log_o.grad) <= GRAD_TOL
    assert H.rel_err(cam_g.grad.cpu(), cam_o.grad) <= GRAD_TOL
    assert H.rel_err(off_g.grad.cpu(), off_o.grad) <= GRAD_TOL
    assert H.rel_err(ref_g.grad.cpu(), ref_o.grad) <= GRAD_TOL
    gtol = GRAD_TOL if dtype == torch.float32 else 1e-2        # grads are cast to bf16 on the way back
    for fg, fo in zip(feats_g, feats_o):
        assert H.rel_err(fg.grad.cpu().float(), fo.grad) <= gtol


def test_pack_is_exact_and_zero_copy_for_channels_last():
    sc = H.scene(B=2, T=1, Q=8)
    f = sc["feats"][0].cuda()
    packed = ops.pack_level(f)
    B, N, C, Hh, W = f.shape
    assert torch.equal(packed, f.flatten(0, 1).permute(0, 2, 3, 1).contiguous())
    f_cl = f.flatten(0, 1).contiguous(memory_format=torch.channels_last).unflatten(0, (B, N))
    p2 = ops.pack_level(f_cl)
    assert p2.data_ptr() == f_cl.data_ptr()
    assert torch.equal(p2, packed)
    pb = ops.pack_level(f, torch.bfloat16)
    assert torch.equal(pb, packed.to(torch.bfloat16))


def test_no_cpu_fallback():
    sc = H.scene(B=1, T=1, Q=8)
    with pytest.raises(RuntimeError):
        ops.pack_features(sc["feats"])          # CPU tensors must be refused, not silently computed


def test_error_codes():
    from graph_detr4d_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    p = _lib.XViewParams()
    assert lib.gd4d_xview_forward(None, None) == -1
    p.abi_version = 99
    assert lib.gd4d_xview_forward(C.byref(p), None) == -5
    p.abi_version = _lib.ABI_VERSION
    assert lib.gd4d_xview_forward(C.byref(p), None) == -2       # all dims zero
    p.B = p.Q = p.N = p.L = p.P = 1
    p.Hh, p.C = 8, 128                                           # head width 16
    assert lib.gd4d_xview_forward(C.byref(p), None) == -3


# ------------------------------------------------------------------------------------------
# wide (gather-then-project) mode
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,dtype,T,Q", [(256, torch.float32, 2, 96), (128, torch.float32, 1, 50),
                                         (256, torch.bfloat16, 2, 64), (512, torch.bfloat16, 1, 20)])
def test_mode_c_wide_forward_backward(C, dtype, T, Q):
    sc = H.scene(B=1, T=T, Q=Q, C=C)
    logits, offsets, cam = H.rand_inputs_c(sc, P=4)
    g = torch.Generator().manual_seed(10)
    g1 = torch.randn(1, 8, Q, C, generator=g)
    g2 = torch.randn(1, 8, Q, generator=g)
    feats_src = [f.to(dtype).float() for f in sc["feats"]]
    feats_o = [_leaf(f) for f in feats_src]
    ref_o, log_o, off_o, cam_o = _leaf(sc["ref"]), _leaf(logits), _leaf(offsets), _leaf(cam)
    agg_o, ws_o = xo.xview_c_wide_core(feats_o, ref_o, off_o, log_o, cam_o, sc["l2i"], syn.PC_RANGE,
                                       900, 1600, 8)
    ((agg_o * g1).sum() + (ws_o * g2).sum()).backward()

    feats_g = [_leaf(f.cuda()) for f in feats_src]
    ref_g, log_g, off_g, cam_g = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
    packed = ops.pack_features(feats_g, dtype)
    assert packed.token is not None and packed.levels[0].dtype == dtype
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    agg, ws = ops.xview_attention(cfg, packed, ref_g, log_g, off_g, cam_g, sc["l2i"].cuda())
    assert tuple(agg.shape) == (1, 8, Q, C) and tuple(ws.shape) == (1, 8, Q)
    ((agg * g1.cuda()).sum() + (ws * g2.cuda()).sum()).backward()
    assert H.rel_err(agg.detach().cpu(), agg_o.detach()) <= FWD_TOL
    assert H.rel_err(ws.detach().cpu(), ws_o.detach()) <= FWD_TOL
    assert H.rel_err(log_g.grad.cpu(), log_o.grad) <= GRAD_TOL
    assert H.rel_err(cam_g.grad.cpu(), cam_o.grad) <= GRAD_TOL
    assert H.rel_err(off_g.grad.cpu(), off_o.grad) <= GRAD_TOL
    assert H.rel_err(ref_g.grad.cpu(), ref_o.grad) <= GRAD_TOL
    for fg, fo in zip(feats_g, feats_o):
        assert fg.grad.dtype == torch.float32
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL


@pytest.mark.parametrize("dtype,T,B,static,C", [(torch.float32, 1, 1, False, 256), (torch.float32, 2, 2, False, 256),
                                                (torch.bfloat16, 2, 1, False, 256), (torch.float32, 1, 1, True, 256),
                                                (torch.bfloat16, 1, 1, False, 512), (torch.float32, 1, 1, False, 128)])
def test_sorted_backward_equals_atomics_backward_and_leaves_its_scratch_clean(dtype, T, B, static, C):
    """The owner-computes backward (xview_bwd_sorted.cu: sort by pixel row, one reduction per run) against
    the atomics backward (xview_bwd.cu): same gradients up to fp32 summation order; calling it three
    times on one scratch gives the same answer (its counters and histogram are left zeroed); with the
    static one-warp-per-item schedule too; NULL feature-gradient maps are skipped."""
    sc = H.scene(B=B, T=T, Q=333, C=C)
    logits, offsets, cam = H.rand_inputs_c(sc)
    packed = _pack(sc["feats"], dtype)
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    args = (cfg, packed.levels, B, sc["N"], sc["ref"].cuda(), logits.cuda(), offsets.cuda(), cam.cuda(),
            sc["l2i"].cuda())
    g = torch.Generator().manual_seed(3)
    go = torch.randn(B, 8, 333, C, generator=g).cuda()
    gw = torch.randn(B, 8, 333, generator=g).cuda()
    res = {}
    ops.DYNAMIC_SCHEDULE = not static
    try:
        for srt in (False, True, True, True):
            ops.SORTED_BACKWARD = srt
            gv = [torch.zeros(v.shape, device="cuda") for v in packed.levels]
            n0 = ops.launch_count()
            small = ops.xview_backward(*args, go, gv, grad_wsum=gw)
            assert ops.launch_count() - n0 == (5 if srt else 1)
            res.setdefault(srt, []).append((gv, small))
        small_only = ops.xview_backward(*args, go, None, grad_wsum=gw)      # no feature gradients wanted
    finally:
        ops.SORTED_BACKWARD, ops.DYNAMIC_SCHEDULE = "auto", True
    gv_a, small_a = res[False][0]
    for gv_s, small_s in res[True]:
        for a, b in zip(gv_s, gv_a):
            assert H.rel_err(a, b) <= 1e-5 and float(b.abs().max()) > 0
        for a, b in zip(small_s, small_a):
            assert H.rel_err(a, b) <= 2e-5
    for a, b in zip(small_only, small_a):
        assert H.rel_err(a, b) <= 2e-5


def test_grad_sink_is_shared_across_layers():
    """Six 'layers' sampling the same packed maps must accumulate into ONE grad map."""
    sc = H.scene(B=1, T=1, Q=40)
    feats_g = [_leaf(f.cuda()) for f in sc["feats"]]
    packed = ops.pack_features(feats_g)
    cfg = XViewConfig(MODE_A, 8, 1, tuple(syn.PC_RANGE), 900.0, 1600.0)
    total = 0
    for i in range(3):
        logits = H.rand_inputs_a(sc, seed=20 + i).cuda()
        total = total + ops.xview_attention(cfg, packed, sc["ref"].cuda(), logits,
                                            lidar2img=sc["l2i"].cuda()).sum()
    total.backward()
    assert packed.sink.buffers is None                       # handed over exactly once
    feats_o = [_leaf(f) for f in sc["feats"]]
    tot_o = 0
    for i in range(3):
        out_o, _ = xo.xview_a_core(feats_o, sc["ref"], H.rand_inputs_a(sc, seed=20 + i), sc["l2i"],
                                   syn.PC_RANGE, 900, 1600)
        tot_o = tot_o + out_o.sum()
    tot_o.backward()
    for fg, fo in zip(feats_g, feats_o):
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL


def test_dynamic_and_static_schedules_agree_and_counter_resets():
    sc = H.scene(B=2, T=2, Q=333)
    logits, offsets, cam = H.rand_inputs_c(sc)
    packed = _pack(sc["feats"])
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    args = (cfg, packed.levels, 2, sc["N"], sc["ref"].cuda(), logits.cuda(), offsets.cuda(), cam.cuda(),
            sc["l2i"].cuda())
    outs = {}
    for dyn in (True, False):
        ops.DYNAMIC_SCHEDULE = dyn
        try:
            for rep in range(3):                    # repeated launches reuse the self-resetting counter
                (o, ws), _ = ops.xview_forward(*args)
            gv = [torch.zeros(v.shape, device="cuda") for v in packed.levels]
            g = ops.xview_backward(*args, torch.ones_like(o), gv, grad_wsum=torch.ones_like(ws))
            outs[dyn] = (o, ws, gv, g)
        finally:
            ops.DYNAMIC_SCHEDULE = True
    torch.cuda.synchronize()
    for t in ops._SCHED.values():
        assert int(t.abs().sum()) == 0              # every launch left the counter at zero
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    for a, b in zip(outs[True][2], outs[False][2]):
        assert H.rel_err(a, b) <= 1e-5              # atomics: order differs, values agree
    assert H.rel_err(outs[True][3][0], outs[False][3][0]) <= 1e-5


# ------------------------------------------------------------------------------------------
# mode V2 (Detr3DCrossAttenV2)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,Q,dtype", [(1, 70, torch.float32), (2, 33, torch.float32), (1, 40, torch.bfloat16)])
def test_mode_v2_forward_backward(T, Q, dtype):
    from graph_detr4d_b200.ops import MODE_V2
    sc = H.scene(B=1, T=T, Q=Q)
    N, Hh, L, P = sc["N"], 8, 4, 4
    g = torch.Generator().manual_seed(12)
    logits = torch.randn(1, Q, N * Hh * L * P, generator=g)
    offsets = torch.randn(1, Q, N * Hh * L * P * 2, generator=g) * 3.0
    gout = torch.randn(1, Q, sc["C"], generator=g)
    feats_src = [f.to(dtype).float() for f in sc["feats"]]
    feats_o = [_leaf(f) for f in feats_src]
    ref_o, log_o, off_o = _leaf(sc["ref"]), _leaf(logits), _leaf(offsets)
    out_o, mask_o = xo.xview_v2_core(feats_o, ref_o, off_o, log_o, sc["l2i"], syn.PC_RANGE, 900, 1600, Hh)
    out_o.backward(gout)
    feats_g = [_leaf(f.cuda()) for f in feats_src]
    ref_g, log_g, off_g = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets))
    packed = ops.pack_features(feats_g, dtype)
    cfg = XViewConfig(MODE_V2, Hh, P, tuple(syn.PC_RANGE), 900.0, 1600.0)
    out = ops.xview_attention(cfg, packed, ref_g, log_g, off_g, None, sc["l2i"].cuda())
    out.backward(gout.cuda())
    _, mask = ops.xview_forward(cfg, packed.levels, 1, N, sc["ref"].cuda(), logits.cuda(), offsets.cuda(),
                                None, sc["l2i"].cuda(), want_mask=True)
    assert torch.equal(mask.cpu().bool(), mask_o)
    assert H.rel_err(out.detach().cpu(), out_o.detach()) <= FWD_TOL
    assert H.rel_err(log_g.grad.cpu(), log_o.grad) <= GRAD_TOL
    assert H.rel_err(off_g.grad.cpu(), off_o.grad) <= GRAD_TOL
    assert H.rel_err(ref_g.grad.cpu(), ref_o.grad) <= GRAD_TOL
    for fg, fo in zip(feats_g, feats_o):
        assert H.rel_err(fg.grad.cpu(), fo.grad) <= GRAD_TOL


@pytest.mark.parametrize("C,dtype,T,Q", [(256, torch.float32, 2, 333), (256, torch.bfloat16, 2, 200),
                                         (128, torch.float32, 1, 64)])
def test_tma_forward_path_matches_register_gather_path(C, dtype, T, Q):
    """cp.async.bulk + mbarrier staging (xview_fwd_tma.cu) vs the LDG kernel and the oracle."""
    sc = H.scene(B=1, T=T, Q=Q, C=C)
    logits, offsets, cam = H.rand_inputs_c(sc)
    feats_src = [f.to(dtype).float() for f in sc["feats"]]
    packed = _pack(feats_src, dtype)
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    args = (cfg, packed.levels, 1, sc["N"], sc["ref"].cuda(), logits.cuda(), offsets.cuda(), cam.cuda(),
            sc["l2i"].cuda())
    (o_ldg, ws_ldg), _ = ops.xview_forward(*args)
    ops.TMA_FORWARD = True
    try:
        for _ in range(3):                              # mbarrier phases must survive relaunches
            (o_tma, ws_tma), m = ops.xview_forward(*args, want_mask=True)
    finally:
        ops.TMA_FORWARD = False
    agg_o, ws_o = xo.xview_c_wide_core(feats_src, sc["ref"], offsets, logits, cam, sc["l2i"], syn.PC_RANGE,
                                       900, 1600, 8)
    assert H.rel_err(o_tma.cpu(), agg_o) <= FWD_TOL and H.rel_err(ws_tma.cpu(), ws_o) <= FWD_TOL
    assert H.rel_err(o_tma, o_ldg) <= 1e-6 and H.rel_err(ws_tma, ws_ldg) <= 1e-6


@pytest.mark.parametrize("C,H_,W_", [(256, 12, 20), (36, 8, 13), (64, 29, 50), (256, 15, 25), (40, 4, 33)])
def test_pack_kernels_exact_all_paths(C, H_, W_):
    """Vectorised swizzled path (HW % 4 == 0) and the scalar fallback, tile tails included."""
    g = torch.Generator().manual_seed(C + H_)
    f = torch.randn(2, 3, C, H_, W_, generator=g).cuda()
    want = f.flatten(0, 1).permute(0, 2, 3, 1).contiguous()
    assert torch.equal(ops.pack_level(f), want)
    assert torch.equal(ops.pack_level(f, torch.bfloat16), want.to(torch.bfloat16))
    fb = f.to(torch.bfloat16)                                                   # bf16 producer: bf16 -> bf16 paths
    assert torch.equal(ops.pack_level(fb), fb.flatten(0, 1).permute(0, 2, 3, 1).contiguous())


def test_bf16_producer_gets_its_bf16_nchw_gradient_from_one_launch():
    """bf16 NCHW leaf maps: the shared fp32 channel-last grad map comes back as a bf16 NCHW tensor through
    gd4d_unpack_nhwc_cast (every level, incl. the ones whose H*W is not a multiple of 4); equal to converting the
    gradient the fp32-leaf path produces."""
    sc = H.scene(B=1, T=1, Q=96)
    logits, offsets, cam = H.rand_inputs_c(sc)
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    g = torch.Generator().manual_seed(9)
    g1 = torch.randn(1, 8, 96, 256, generator=g).cuda()
    grads = {}
    for leaf_dtype in (torch.float32, torch.bfloat16):
        feats = [_leaf(f.to(torch.bfloat16).to(leaf_dtype).cuda()) for f in sc["feats"]]
        packed = ops.pack_features(feats, torch.bfloat16)
        agg, ws = ops.xview_attention(cfg, packed, sc["ref"].cuda(), logits.cuda(), offsets.cuda(), cam.cuda(),
                                      sc["l2i"].cuda())
        (agg * g1).sum().backward()
        grads[leaf_dtype] = [f.grad for f in feats]
    for a, b in zip(grads[torch.bfloat16], grads[torch.float32]):
        assert a.dtype == torch.bfloat16 and a.is_contiguous() and a.shape == b.shape
        # fp32 atomics order differs run to run by ~1e-7: compare after the same rounding, 1 bf16 ulp
        assert H.rel_err(a.float(), b.to(torch.bfloat16).float()) <= 2 ** -7


def test_presorted_backward_on_a_side_stream_matches():
    """gd4d_xview_backward_sort right after the forward on a side stream + GD4D_FLAG_BWD_PRESORTED backward
    (ops.PRESORT, opt-in) gives the gradients of the one-call sorted backward."""
    sc = H.scene(B=1, T=2, Q=200)
    logits, offsets, cam = H.rand_inputs_c(sc)
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    g = torch.Generator().manual_seed(4)
    g1, g2 = torch.randn(1, 8, 200, 256, generator=g).cuda(), torch.randn(1, 8, 200, generator=g).cuda()
    res = {}
    try:
        for pre in (False, True):
            ops.PRESORT, ops.SORTED_BACKWARD = pre, True
            feats = [_leaf(f.cuda()) for f in sc["feats"]]
            ref, log, off, cm = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
            packed = ops.pack_features(feats)
            n0 = ops.launch_count()
            agg, ws = ops.xview_attention(cfg, packed, ref, log, off, cm, sc["l2i"].cuda())
            assert ops.launch_count() - n0 == (4 if pre else 1)            # forward (+ the 3 sort kernels)
            ((agg * g1).sum() + (ws * g2).sum()).backward()
            res[pre] = [t.grad.clone() for t in (ref, log, off, cm, *feats)]
    finally:
        ops.PRESORT, ops.SORTED_BACKWARD = False, "auto"
    for a, b in zip(res[True], res[False]):
        assert H.rel_err(a, b) <= 1e-5 and float(b.abs().max()) > 0


@pytest.mark.parametrize("dtype,T", [(torch.float32, 1), (torch.float32, 2), (torch.bfloat16, 2)])
def test_forward_emitted_records_give_the_same_backward(dtype, T):
    """GD4D_FLAG_FWD_EMIT: the forward kernel writes the sorted backward's contribution records itself; the backward
    (scan, scatter, owner, finish: 4 launches) must equal the five-launch sorted backward and the oracle-checked
    atomics backward; the forward outputs are untouched by the emission."""
    sc = H.scene(B=1, T=T, Q=200)
    logits, offsets, cam = H.rand_inputs_c(sc)
    cfg = XViewConfig(MODE_C, 8, 4, tuple(syn.PC_RANGE), 900.0, 1600.0, wide=True)
    g = torch.Generator().manual_seed(4)
    g1, g2 = torch.randn(1, 8, 200, 256, generator=g).cuda(), torch.randn(1, 8, 200, generator=g).cuda()
    res = {}
    try:
        for name, srt, emit in (("atomics", False, False), ("sorted", True, False), ("emitted", True, True)):
            ops.SORTED_BACKWARD, ops.FWD_EMIT, ops.PRESORT = srt, emit, False
            feats = [_leaf(f.cuda()) for f in sc["feats"]]
            ref, log, off, cm = (_leaf(t.cuda()) for t in (sc["ref"], logits, offsets, cam))
            packed = ops.pack_features(feats, dtype)
            agg, ws = ops.xview_attention(cfg, packed, ref, log, off, cm, sc["l2i"].cuda())
            n0 = ops.launch_count()
            ((agg * g1).sum() + (ws * g2).sum()).backward()
            n_bwd = ops.launch_count() - n0
            res[name] = ([agg.detach().clone(), ws.detach().clone()], [t.grad.clone() for t in (ref, log, off, cm, *feats)], n_bwd)
    finally:
        ops.SORTED_BACKWARD, ops.FWD_EMIT, ops.PRESORT = "auto", True, False
    assert res["emitted"][2] < res["sorted"][2]                              # one kernel fewer in the backward
    for a, b in zip(res["emitted"][0], res["atomics"][0]):
        assert torch.equal(a, b)                                             # same forward, bit for bit
    for key in ("sorted", "emitted"):
        for a, b in zip(res[key][1], res["atomics"][1]):
            assert H.rel_err(a, b) <= 2e-5 and float(b.abs().max()) > 0
