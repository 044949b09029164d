"""The C-ABI library loads without a GPU and exports every symbol the header
declares; parameter validation (which runs before any launch) returns the
documented status codes."""
import ctypes as C
import os
import re

from graph_detr4d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    inc = os.path.join(ROOT, "include")
    src = "".join(open(os.path.join(inc, h)).read() for h in sorted(os.listdir(inc)) if h.endswith(".h"))
    return re.findall(r"GD4D_API\s+[\w\s\*]+?\b(gd4d_\w+)\s*\(", src)


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = _declared()
    assert set(names) == set(_lib.EXPORTS) and len(names) >= 7
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layout_and_version():
    lib = _lib.load()
    assert lib.gd4d_abi_version() == _lib.ABI_VERSION
    assert lib.gd4d_params_size() == C.sizeof(_lib.XViewParams)


def _valid_params():
    p = _lib.XViewParams()
    p.abi_version = _lib.ABI_VERSION
    p.mode, p.value_dtype = _lib.MODE_C, _lib.F32
    p.B, p.Q, p.N, p.Hh, p.L, p.P, p.C = 1, 900, 12, 8, 4, 4, 256
    for l, (h, w) in enumerate([(116, 200), (58, 100), (29, 50), (15, 25)]):
        p.level_h[l], p.level_w[l] = h, w
        p.value[l] = 0x1000 * (l + 1)            # fake, aligned, never dereferenced on the host
    p.img_h, p.img_w = 900.0, 1600.0
    p.ref = p.lidar2img = p.attn_logits = p.offsets = p.cam_logits = p.out = 0x10000
    return p


def test_launch_info_and_status_codes():
    lib = _lib.load()
    p = _valid_params()
    g, b, s = C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.gd4d_xview_launch_info(C.byref(p), C.byref(g), C.byref(b), C.byref(s)) == 0
    assert b.value == 256 and g.value == 900 and s.value > 0
    assert lib.gd4d_xview_launch_info(None, None, None, None) == -1
    q = _valid_params(); q.abi_version = 7
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -5
    q = _valid_params(); q.Q = 0
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -2
    q = _valid_params(); q.C = 512
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -3
    q = _valid_params(); q.value[2] = 0x1004
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -4
    q = _valid_params(); q.value[1] = None
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -1
    q = _valid_params(); q.offsets = None
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -1
    q = _valid_params(); q.P = 32                                   # L*P = 128 > 64
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -5
    q = _valid_params(); q.mode = 7
    assert lib.gd4d_xview_launch_info(C.byref(q), None, None, None) == -5
    q = _valid_params(); q.grad_out = None
    assert lib.gd4d_xview_backward(C.byref(q), None) == -1          # validation precedes any launch
    assert lib.gd4d_pack_nchw(None, None, 0, 0, 1, 1, 1, 1, None) == -1
    assert lib.gd4d_pack_nchw(0x1000, 0x2000, 0, 0, 0, 1, 1, 1, None) == -2
    # glue entry points validate before launching, too
    assert lib.gd4d_inverse_sigmoid_fwd(None, None, 1, 1e-5, 0, None) == -1
    assert lib.gd4d_inverse_sigmoid_bwd(0x1000, 0x1000, 0x1000, 0, 1e-5, 0, None) == -2
    assert lib.gd4d_ref_update(0x1000, 4, 0x1000, 0x1000, 10, 1e-5, None) == -2      # reg_stride < 5
    a = 0x1000
    ln = lib.gd4d_add_layernorm_fwd          # x, xbias, r1, r2, gamma, beta, pos, y, y2, y_copy1, y_copy2, s_out, mean, rstd, ...
    assert ln(a, None, None, None, a, a, None, a, None, None, None, None, a, a, 900, 200, 1e-5, 0, None) == -2
    assert ln(a + 4, None, None, None, a, a, None, a, None, None, None, None, a, a, 900, 256, 1e-5, 0, None) == -4
    assert ln(a, a, None, None, a, a, None, a, None, None, None, None, a, a, 900, 256, 1e-5, 0, None) == -1    # s_out
    assert ln(a, None, None, None, a, a, a, a, None, None, None, None, a, a, 900, 256, 1e-5, 0, None) == -1    # pos w/o y2
    assert ln(a, None, None, None, a, a, None, a, None, a + 4, None, None, a, a, 900, 256, 1e-5, 0, None) == -4  # copy alignment
    assert lib.gd4d_bias_act(a, a, 900, 10, 1, None) == -2                                      # C % 4
    assert lib.gd4d_add_layernorm_bwd(a, None, None, None, a, a, a, a, None, a, None, 900, 256, 1, None) == -1   # relu needs beta
    assert lib.gd4d_add_layernorm_bwd(None, None, None, None, a, a, a, a, a, a, None, 900, 256, 0, None) == -1   # no gradient at all
    assert lib.gd4d_unpack_nhwc(a, None, 1, 1, 1, 1, None) == -1
    assert lib.gd4d_match_cost(a, a, a, a, a, 900, 10, 7, 5, 9, 2.0, 0.25, 0.25, 1e-12, None) == -2   # code < 8
    pc = (C.c_float * 6)(*[0.0] * 6)
    assert lib.gd4d_frustum_pe(a, None, None, None, 6, 4, 4, 8, 928.0, 1600.0, 1.0, 0.1, pc, None) == -1
    assert lib.gd4d_frustum_pe(a, None, a, None, 6, 4, 4, 0, 928.0, 1600.0, 1.0, 0.1, pc, None) == -2
    # r2 entry points: validation only, nothing launches without a GPU
    ptrs = (C.c_void_p * 1)(a)
    hs, ws_ = (C.c_int32 * 1)(4), (C.c_int32 * 1)(4)
    assert lib.gd4d_frustum_pe_levels(a, None, ptrs, None, 6, 9, hs, ws_, 8, 928.0, 1600.0, 1.0, 0.1, pc, None) == -2   # > 8 levels
    assert lib.gd4d_frustum_pe_levels(a, None, None, None, 6, 1, hs, ws_, 8, 928.0, 1600.0, 1.0, 0.1, pc, None) == -1
    assert lib.gd4d_level_mask(None, a, 6, 4, 4, 928, 1600, None) == -1
    assert lib.gd4d_sine_pe3d(a, a, a, 1, 6, 4, 4, 928, 1600, 7, 1, 6.28, 1e-6, -0.5, None) == -2          # odd num_feats
    assert lib.gd4d_fpe_combine_fwd(a, a, a, a, a, 0, None) == -2
    assert lib.gd4d_softmax_bwd(a, a, a, 10, 902, None) == -2                                          # cols % 4
    assert lib.gd4d_softmax_bwd(a, a, a, 10, 2048, None) == -2                                         # cols > 1024
    for fn in (lib.gd4d_gemm_tf32x3, lib.gd4d_sgemm_small):
        assert fn(None, 256, 0, a, 256, 0, a, 256, None, 0, 900, 256, 256, 1, 0, 0, 0, None) == -1
        assert fn(a, 256, 0, a, 256, 0, a, 256, None, 0, 900, 256, 0, 1, 0, 0, 0, None) == -2              # K = 0
        assert fn(a + 4, 256, 0, a, 256, 0, a, 256, None, 0, 900, 256, 256, 1, 0, 0, 0, None) == -4        # alignment
        assert fn(a, 254, 0, a, 256, 0, a, 256, None, 0, 900, 256, 256, 1, 0, 0, 0, None) == -4            # lda % 4
        assert fn(a, 256, 0, a, 256, 0, a, 256, None, 0, 900, 256, 250, 1, 0, 0, 0, None) == -5            # K % 4
    q = _valid_params()
    assert lib.gd4d_xview_bwd_ws_bytes(C.byref(q)) == -5                                               # not wide mode C
    q = _valid_params(); q.mode, q.wide = _lib.MODE_C, 1
    need = lib.gd4d_xview_bwd_ws_bytes(C.byref(q))
    assert need > 0 and need % 256 == 0
    q.bwd_ws, q.bwd_ws_bytes = 0x10000, need - 256
    assert lib.gd4d_xview_backward_sort(C.byref(q), None) == -2                                        # scratch too small
    q.bwd_ws = 0x10010
    assert lib.gd4d_xview_backward_sort(C.byref(q), None) == -4                                        # 256-byte alignment
    q.bwd_ws, q.flags = None, _lib.FLAG_BWD_PRESORTED
    q.grad_out = a
    assert lib.gd4d_xview_backward(C.byref(q), None) == -5                                             # presorted without a scratch
    for code in (0, -1, -2, -3, -4, -5, -6, -99):
        assert len(_lib.strerror(code)) > 0
