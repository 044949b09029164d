"""ctypes binding of libgd4d_xview.so (the C ABI in include/gd4d_xview.h, gd4d_glue.h, gd4d_frustum.h and gd4d_assign.h).

There is deliberately NO fallback: if the shared library is missing or was not
built for this GPU, every op raises.  ``load()`` builds in-tree with nvcc when the
library is absent (the build box has nvcc but no GPU; the GPU box receives the
prebuilt .so with the repo snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

ABI_VERSION = 6
MAX_LEVELS = 8
MODE_A, MODE_C, MODE_V2 = 0, 1, 2
F32, BF16 = 0, 1
FLAG_TMA_FORWARD = 1
FLAG_L2_PREFETCH = 2
FLAG_BWD_SKIP_OWNER = 4        # measurement only (include/gd4d_xview.h)
FLAG_BWD_PRESORTED = 8
FLAG_FWD_EMIT = 16
FLAG_BWD_EMITTED = 32

EXPORTS = (
    "gd4d_abi_version",
    "gd4d_params_size",
    "gd4d_strerror",
    "gd4d_xview_launch_info",
    "gd4d_xview_forward",
    "gd4d_xview_backward",
    "gd4d_xview_bwd_ws_bytes",
    "gd4d_xview_backward_sort",
    "gd4d_pack_nchw",
    "gd4d_unpack_nhwc",
    "gd4d_unpack_nhwc_cast",
    # include/gd4d_glue.h
    "gd4d_inverse_sigmoid_fwd",
    "gd4d_inverse_sigmoid_bwd",
    "gd4d_ref_update",
    "gd4d_bias_act",
    "gd4d_add_layernorm_fwd",
    "gd4d_add_layernorm_bwd",
    "gd4d_adamw_chunk",
    "gd4d_adamw_multi",
    "gd4d_softmax_bwd",
    "gd4d_gemm_tf32x3",
    "gd4d_sgemm_small",
    # include/gd4d_frustum.h
    "gd4d_frustum_pe",
    "gd4d_frustum_pe_levels",
    # include/gd4d_assign.h
    "gd4d_match_cost",
    # include/gd4d_fpe.h
    "gd4d_level_mask",
    "gd4d_sine_pe3d",
    "gd4d_fpe_combine_fwd",
    "gd4d_fpe_combine_bwd",
)


class XViewParams(C.Structure):
    """Mirror of ``gd4d_xview_params`` (include/gd4d_xview.h) -- keep in sync."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("mode", C.c_int32),
        ("value_dtype", C.c_int32),
        ("B", C.c_int32), ("Q", C.c_int32), ("N", C.c_int32), ("Hh", C.c_int32),
        ("L", C.c_int32), ("P", C.c_int32), ("C", C.c_int32),
        ("wide", C.c_int32),
        ("flags", C.c_uint32),
        ("gen_stride", C.c_int32),
        ("level_h", C.c_int32 * MAX_LEVELS),
        ("level_w", C.c_int32 * MAX_LEVELS),
        ("pc_lo", C.c_float * 3),
        ("pc_span", C.c_float * 3),
        ("img_h", C.c_float), ("img_w", C.c_float),
        ("value", C.c_void_p * MAX_LEVELS),
        ("value_bias", C.c_void_p),
        ("ref", C.c_void_p),
        ("lidar2img", C.c_void_p),
        ("attn_logits", C.c_void_p),
        ("offsets", C.c_void_p),
        ("cam_logits", C.c_void_p),
        ("sched", C.c_void_p),
        ("out", C.c_void_p),
        ("wsum", C.c_void_p),
        ("mask", C.c_void_p),
        ("grad_out", C.c_void_p),
        ("grad_wsum", C.c_void_p),
        ("grad_value", C.c_void_p * MAX_LEVELS),
        ("grad_value_bias", C.c_void_p),
        ("grad_attn_logits", C.c_void_p),
        ("grad_offsets", C.c_void_p),
        ("grad_cam_logits", C.c_void_p),
        ("grad_ref", C.c_void_p),
        ("bwd_ws", C.c_void_p),
        ("bwd_ws_bytes", C.c_int64),
    ]


class Gd4dError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what}: gd4d status {status} ({strerror(status)})")


_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if needed) and type the C-ABI library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
            _build.build()
        lib = C.CDLL(path)
        lib.gd4d_abi_version.restype = C.c_int
        lib.gd4d_abi_version.argtypes = []
        lib.gd4d_strerror.restype = C.c_char_p
        lib.gd4d_strerror.argtypes = [C.c_int]
        lib.gd4d_xview_launch_info.restype = C.c_int
        lib.gd4d_xview_launch_info.argtypes = [C.POINTER(XViewParams), C.POINTER(C.c_int32),
                                               C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.gd4d_xview_forward.restype = C.c_int
        lib.gd4d_xview_forward.argtypes = [C.POINTER(XViewParams), C.c_void_p]
        lib.gd4d_xview_backward.restype = C.c_int
        lib.gd4d_xview_backward.argtypes = [C.POINTER(XViewParams), C.c_void_p]
        lib.gd4d_xview_backward_sort.restype = C.c_int
        lib.gd4d_xview_backward_sort.argtypes = [C.POINTER(XViewParams), C.c_void_p]
        lib.gd4d_xview_bwd_ws_bytes.restype = C.c_int64
        lib.gd4d_xview_bwd_ws_bytes.argtypes = [C.POINTER(XViewParams)]
        lib.gd4d_pack_nchw.restype = C.c_int
        lib.gd4d_pack_nchw.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.gd4d_unpack_nhwc_cast.restype = C.c_int
        lib.gd4d_unpack_nhwc_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_void_p]
        lib.gd4d_unpack_nhwc.restype = C.c_int
        lib.gd4d_unpack_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p]
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        for name, args in (
                ("gd4d_inverse_sigmoid_fwd", [vp, vp, i64, f32, i32, vp]),
                ("gd4d_inverse_sigmoid_bwd", [vp, vp, vp, i64, f32, i32, vp]),
                ("gd4d_ref_update", [vp, i32, vp, vp, i64, f32, vp]),
                ("gd4d_bias_act", [vp, vp, i64, i32, i32, vp]),
                ("gd4d_add_layernorm_fwd", [vp] * 14 + [i64, i32, f32, i32, vp]),
                ("gd4d_add_layernorm_bwd", [vp] * 11 + [i64, i32, i32, vp]),
                ("gd4d_adamw_chunk", []),
                ("gd4d_adamw_multi", [vp, vp, i32, vp, f32, f32, f32, f32, f32, vp]),
                ("gd4d_softmax_bwd", [vp, vp, vp, i64, i32, vp]),
                ("gd4d_gemm_tf32x3", [vp, i64, i32, vp, i64, i32, vp, i64, vp, i32, i32, i32, i32, i32, i64, i64, i64, vp]),
                ("gd4d_sgemm_small", [vp, i64, i32, vp, i64, i32, vp, i64, vp, i32, i32, i32, i32, i32, i64, i64, i64, vp]),
                ("gd4d_frustum_pe", [vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, f32, f32,
                                     C.POINTER(C.c_float), vp]),
                ("gd4d_frustum_pe_levels", [vp, vp, vp, vp, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i32,
                                            f32, f32, f32, f32, C.POINTER(C.c_float), vp]),
                ("gd4d_match_cost", [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, f32, f32, f32, vp]),
                ("gd4d_level_mask", [vp, vp, i32, i32, i32, i32, i32, vp]),
                ("gd4d_sine_pe3d", [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32, f32, f32, vp]),
                ("gd4d_fpe_combine_fwd", [vp, vp, vp, vp, vp, i64, vp]),
                ("gd4d_fpe_combine_bwd", [vp, vp, vp, vp, vp, i64, vp])):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = C.c_int, args
        lib.gd4d_params_size.restype = C.c_int
        lib.gd4d_params_size.argtypes = []
        if lib.gd4d_abi_version() != ABI_VERSION:
            raise RuntimeError("libgd4d_xview.so ABI version mismatch; rebuild")
        if lib.gd4d_params_size() != C.sizeof(XViewParams):
            raise RuntimeError("gd4d_xview_params layout mismatch between the header and _lib.XViewParams")
        _lib = lib
    return _lib


def strerror(status: int) -> str:
    return load().gd4d_strerror(int(status)).decode()


def check(status: int, what: str):
    if status != 0:
        raise Gd4dError(status, what)
