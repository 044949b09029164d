"""Host side of the fp32-accurate tensor-core GEMM (include/gd4d_glue.h ``gd4d_gemm_tf32x3``,
csrc/gemm_tf32x3.cu): torch tensors in, one C-ABI launch out.  CUDA only, no fallback inside:
``supported`` tells the caller (glue.py) whether a shape can take this path; otherwise the caller keeps
its library GEMM."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .ops import _count, _stream_ptr

__all__ = ["gemm", "supported"]


def _row_major_2d(t: torch.Tensor):
    """(data_ptr-compatible tensor, leading dimension) if ``t`` (2-D or batched 3-D) has unit stride in
    its last dim and a row stride that is a multiple of 4 floats, else None."""
    if t.dtype != torch.float32 or not t.is_cuda or t.stride(-1) != 1:
        return None
    ld = t.stride(-2)
    if ld % 4 != 0 or ld < t.shape[-1] or t.data_ptr() % 16 != 0:
        return None
    return ld


def supported(a: torch.Tensor, b: torch.Tensor, a_t: bool = False, b_t: bool = False) -> bool:
    """``a_t``: A is given as (K,M); ``b_t``: B is given as (K,N)  (see ``gemm``)."""
    if a.dim() != b.dim() or a.dim() not in (2, 3):
        return False
    if _row_major_2d(a) is None or _row_major_2d(b) is None:
        return False
    if a.shape[-1] % 4 != 0 or b.shape[-1] % 4 != 0:           # contiguous extents: 16-byte chunks
        return False
    if a.dim() == 3 and (a.stride(0) % 4 != 0 or b.stride(0) % 4 != 0 or a.shape[0] != b.shape[0]):
        return False
    ka = a.shape[-2] if a_t else a.shape[-1]
    kb = b.shape[-2] if b_t else b.shape[-1]
    return ka == kb and ka > 0


def gemm(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
         a_t: bool = False, b_t: bool = False, out: Optional[torch.Tensor] = None, impl: str = "auto") -> torch.Tensor:
    """C = op(A) . op(B)^T (+ bias) (relu), fp32.  ``impl``: "simt" = exact-fp32 register-tiled FFMA kernel
    (csrc/sgemm_small.cu), "tf32x3" = error-compensated 3xTF32 on the tcgen05 tensor cores
    (csrc/gemm_tf32x3.cu), "auto" = tensor cores once the problem has enough 128-row tiles to fill the
    machine (a tcgen05.mma.kind::tf32 costs 153 cycles whatever its N: at M = 900 the SIMT kernel wins).

    a: (M,K), or (K,M) when ``a_t``;  b: (N,K), or (K,N) when ``b_t``;  optionally batched (leading dim).
    So  x @ W.T -> gemm(x, W);   dY @ W -> gemm(dY, W, b_t=True);   dY.T @ X -> gemm(dY, X, a_t=True, b_t=True)."""
    if not supported(a, b, a_t, b_t):
        raise ValueError("gemm: operands must be CUDA fp32, unit stride in the last dim, 16-byte aligned rows "
                         "(check gemm.supported first)")
    batch = a.shape[0] if a.dim() == 3 else 1
    M = a.shape[-1] if a_t else a.shape[-2]
    K = a.shape[-2] if a_t else a.shape[-1]
    N = b.shape[-1] if b_t else b.shape[-2]
    if out is None:
        out = torch.empty((batch, M, N) if a.dim() == 3 else (M, N), device=a.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or out.stride(-1) != 1 or tuple(out.shape[-2:]) != (M, N):
        raise ValueError("gemm: out must be fp32 (.., M, N) with unit stride in the last dim")
    if bias is not None:
        if bias.dtype != torch.float32 or not bias.is_contiguous() or bias.numel() != N:
            raise ValueError("gemm: bias must be a contiguous fp32 vector of N elements")
    sa = a.stride(0) if a.dim() == 3 else 0
    sb = b.stride(0) if b.dim() == 3 else 0
    sc = out.stride(0) if out.dim() == 3 else 0
    if impl == "auto":
        impl = "tf32x3" if batch * ((M + 127) // 128) * ((N + 63) // 64) >= 256 and K >= 64 else "simt"
    if impl not in ("simt", "tf32x3"):
        raise ValueError(f"gemm: unknown impl {impl!r}")
    lib = _lib.load()
    fn = lib.gd4d_gemm_tf32x3 if impl == "tf32x3" else lib.gd4d_sgemm_small
    st = fn(a.data_ptr(), a.stride(-2), 1 if a_t else 0, b.data_ptr(), b.stride(-2),
                                      1 if b_t else 0, out.data_ptr(), out.stride(-2),
                                      None if bias is None else bias.data_ptr(), 1 if relu else 0, M, N, K, batch,
                                      sa, sb, sc, _stream_ptr(a.device))
    _lib.check(st, "gd4d_gemm_tf32x3" if impl == "tf32x3" else "gd4d_sgemm_small")
    _count()
    return out
