"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/c/xview_ref.c (plain-C restatement of
the forward, reference NCHW layout).  Built by ``__graft_entry__.build()`` / ``oracle/c/Makefile``."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "_build", "libxview_ref.so")


class _P(C.Structure):
    _fields_ = [("mode", C.c_int), ("B", C.c_int), ("Q", C.c_int), ("N", C.c_int), ("Hh", C.c_int),
                ("L", C.c_int), ("P", C.c_int), ("C", C.c_int),
                ("level_h", C.c_int * 8), ("level_w", C.c_int * 8), ("value", C.c_void_p * 8),
                ("ref", C.c_void_p), ("lidar2img", C.c_void_p), ("attn_logits", C.c_void_p),
                ("offsets", C.c_void_p), ("cam_logits", C.c_void_p),
                ("pc_lo", C.c_float * 3), ("pc_span", C.c_float * 3), ("img_h", C.c_float), ("img_w", C.c_float),
                ("out", C.c_void_p), ("mask", C.c_void_p)]


LIB_GLUE = os.path.join(HERE, "c", "_build", "libglue_ref.so")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c"), "all"], check=True)
    return LIB


def pe_frustum(img2lidar, mask_in, H, W, D, pad_h, pad_w, depth_start, bin_size, pc_range):
    """oracle/c/glue_ref.c::pe_frustum -> (out (BN,3D,H,W) float32, mask (BN,H,W) uint8)."""
    if not os.path.exists(LIB_GLUE):
        build()
    lib = C.CDLL(LIB_GLUE)
    m = np.ascontiguousarray(img2lidar, dtype=np.float32).reshape(-1, 16)
    BN = m.shape[0]
    mi = None if mask_in is None else np.ascontiguousarray(mask_in, dtype=np.uint8).reshape(BN, H, W)
    out = np.zeros((BN, 3 * D, H, W), dtype=np.float32)
    mask = np.zeros((BN, H, W), dtype=np.uint8)
    lo = (C.c_float * 3)(*[pc_range[i] for i in range(3)])
    span = (C.c_float * 3)(*[pc_range[3 + i] - pc_range[i] for i in range(3)])
    lib.pe_frustum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    st = lib.pe_frustum(m.ctypes.data, None if mi is None else mi.ctypes.data, out.ctypes.data, mask.ctypes.data,
                        BN, H, W, D, pad_h, pad_w, depth_start, bin_size, lo, span)
    assert st == 0
    return out, mask


def match_cost(cls_pred, bbox_pred, gt, labels, cls_w=2.0, reg_w=0.25, alpha=0.25, eps=1e-12):
    """oracle/c/glue_ref.c::match_cost -> (Q,G) float32."""
    if not os.path.exists(LIB_GLUE):
        build()
    lib = C.CDLL(LIB_GLUE)
    cp = np.ascontiguousarray(cls_pred, dtype=np.float32)
    bp = np.ascontiguousarray(bbox_pred, dtype=np.float32)
    g = np.ascontiguousarray(gt, dtype=np.float32)
    lab = np.ascontiguousarray(labels, dtype=np.int64)
    cost = np.zeros((cp.shape[0], g.shape[0]), dtype=np.float32)
    lib.match_cost.argtypes = [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_float] * 4
    st = lib.match_cost(cp.ctypes.data, bp.ctypes.data, g.ctypes.data, lab.ctypes.data, cost.ctypes.data,
                        cp.shape[0], cp.shape[1], bp.shape[1], g.shape[0], g.shape[1], cls_w, reg_w, alpha, eps)
    assert st == 0
    return cost


def forward(mode, feats, ref, attn_logits, lidar2img, pc_range, img_h, img_w, num_heads, num_points,
            offsets=None, cam_logits=None):
    """feats: list of float32 numpy (B,N,C,H,W); returns out (B,Q,C), mask uint8."""
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    keep = [np.ascontiguousarray(f, dtype=np.float32) for f in feats]
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    a = np.ascontiguousarray(attn_logits, dtype=np.float32)
    m = np.ascontiguousarray(lidar2img, dtype=np.float32)
    B, N, Cc = keep[0].shape[:3]
    Q = ref.shape[1]
    p = _P()
    p.mode, p.B, p.Q, p.N, p.L, p.P, p.C = mode, B, Q, N, len(keep), num_points, Cc
    p.Hh = num_heads if mode == 1 else 1
    for l, f in enumerate(keep):
        p.level_h[l], p.level_w[l] = f.shape[3], f.shape[4]
        p.value[l] = f.ctypes.data
    p.ref, p.lidar2img, p.attn_logits = ref.ctypes.data, m.ctypes.data, a.ctypes.data
    if mode == 1:
        o = np.ascontiguousarray(offsets, dtype=np.float32)
        c = np.ascontiguousarray(cam_logits, dtype=np.float32)
        p.offsets, p.cam_logits = o.ctypes.data, c.ctypes.data
        keep += [o, c]
    for i in range(3):
        p.pc_lo[i] = pc_range[i]
        p.pc_span[i] = pc_range[3 + i] - pc_range[i]
    p.img_h, p.img_w = img_h, img_w
    out = np.zeros((B, Q, Cc), dtype=np.float32)
    mask = np.zeros((B, Q, N) if mode == 0 else (B, N, Q, num_heads, num_points), dtype=np.uint8)
    p.out, p.mask = out.ctypes.data, mask.ctypes.data
    st = lib.xref_forward(C.byref(p))
    if st != 0:
        raise RuntimeError(f"xref_forward status {st}")
    return out, mask
