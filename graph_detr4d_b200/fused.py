"""Host side of the fused per-layer glue kernels (include/gd4d_glue.h, csrc/glue.cu;
SURVEY.md 8f row f2): each call below is ONE launch where the reference (and eager
torch) runs a chain of 5-20 launch-latency-bound one-line ops on 900-row tensors.

  inverse_sigmoid    detr3d_transformer.py:28-43 / deform3d_cross_attn.py:16-31
  ref_update         detr3d_transformer.py:201-214 (result is detached there, so forward-only)
  add_layernorm      post-norm residual sums of the decoder layer and position_encoder's
                     Linear-LN-ReLU stages

CUDA fp32 only, like everything on the path: no eager fallback inside these functions
(the callers in modules.py / decoder.py choose them only for CUDA fp32 tensors).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from .glue import DeferredWgrad
from .ops import _count, _require_cuda, _stream_ptr


# A/B switch for measurements (tools/, bench): GD4D_FUSED_GLUE=0 runs the eager op chains instead.
ENABLED = os.environ.get("GD4D_FUSED_GLUE", "1") != "0"


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


class _InverseSigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps: float, clamp_max: bool):
        xc = _f32c(x, "x")
        y = torch.empty_like(xc)
        st = _lib.load().gd4d_inverse_sigmoid_fwd(xc.data_ptr(), y.data_ptr(), xc.numel(), eps,
                                                  int(clamp_max), _stream_ptr(xc.device))
        _lib.check(st, "gd4d_inverse_sigmoid_fwd")
        _count()
        ctx.save_for_backward(xc)
        ctx.eps, ctx.clamp_max = eps, clamp_max
        return y

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        gy = _f32c(gy, "grad")
        gx = torch.empty_like(xc)
        st = _lib.load().gd4d_inverse_sigmoid_bwd(xc.data_ptr(), gy.data_ptr(), gx.data_ptr(), xc.numel(),
                                                  ctx.eps, int(ctx.clamp_max), _stream_ptr(xc.device))
        _lib.check(st, "gd4d_inverse_sigmoid_bwd")
        _count()
        return gx, None, None


def inverse_sigmoid(x: torch.Tensor, eps: float = 1e-5, clamp_max: bool = False) -> torch.Tensor:
    """One-launch inverse_sigmoid (forward and backward)."""
    if x.numel() == 0:
        return x.clone()
    return _InverseSigmoidFn.apply(x, float(eps), bool(clamp_max))


@torch.no_grad()
def ref_update(reg: torch.Tensor, ref: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """sigmoid(reg[..., (0,1,4)] + inverse_sigmoid(ref)), detached (detr3d_transformer.py:201-214).
    ``reg`` (..., >=5) regression-branch output, ``ref`` (..., 3)."""
    reg = _f32c(reg.detach(), "reg")
    ref = _f32c(ref.detach(), "reference_points")
    if ref.shape[-1] != 3 or reg.shape[:-1] != ref.shape[:-1]:
        raise ValueError(f"ref_update: reg {tuple(reg.shape)} vs ref {tuple(ref.shape)}")
    out = torch.empty_like(ref)
    st = _lib.load().gd4d_ref_update(reg.data_ptr(), int(reg.shape[-1]), ref.data_ptr(), out.data_ptr(),
                                     ref.numel() // 3, float(eps), _stream_ptr(ref.device))
    _lib.check(st, "gd4d_ref_update")
    _count()
    return out


def bias_act_(y: torch.Tensor, bias: torch.Tensor, relu: bool = True) -> torch.Tensor:
    """In place y = [relu](y + bias) over the last dim, one launch (no autograd: used inside
    autograd Functions)."""
    _require_cuda(y, "y")
    if y.dtype != torch.float32 or not y.is_contiguous() or bias.dtype != torch.float32:
        raise TypeError("bias_act_ needs contiguous float32 tensors")
    Cc = y.shape[-1]
    st = _lib.load().gd4d_bias_act(y.data_ptr(), bias.data_ptr(), y.numel() // Cc, Cc, int(relu),
                                   _stream_ptr(y.device))
    _lib.check(st, "gd4d_bias_act")
    _count()
    return y


def softmax_bwd_(grad: torch.Tensor, probs: torch.Tensor) -> torch.Tensor:
    """In place over ``grad``: softmax backward along the last dim, one launch, one pass (ATen: g * p as its
    own elementwise pass, then its kernel).  Contiguous fp32, last dim % 4 == 0 and <= 1024."""
    _require_cuda(grad, "grad")
    cols = grad.shape[-1]
    st = _lib.load().gd4d_softmax_bwd(grad.data_ptr(), probs.data_ptr(), grad.data_ptr(), grad.numel() // cols, cols,
                                      _stream_ptr(grad.device))
    _lib.check(st, "gd4d_softmax_bwd")
    _count()
    return grad


SOFTMAX_BWD = True     # A/B switch (tools/ab_step.py)
LN_COPIES = True       # A/B switch: one output copy per consumer of a fused LayerNorm result (decoder.DecoderLayer)


def can_fuse_softmax_bwd(grad: torch.Tensor, probs: torch.Tensor) -> bool:
    return (ENABLED and SOFTMAX_BWD and grad.is_cuda and grad.dtype == probs.dtype == torch.float32 and grad.is_contiguous()
            and probs.is_contiguous() and grad.shape == probs.shape and grad.shape[-1] % 4 == 0
            and grad.shape[-1] <= 1024 and grad.data_ptr() % 16 == 0 and probs.data_ptr() % 16 == 0)


class _AddLayerNormFn(torch.autograd.Function):
    """y = [relu](LayerNorm(x + xbias + r1 + r2)); gamma/beta gradients go through DeferredWgrad
    when it is active (one batched reduction per step), else they are reduced here.  ``xbias``
    is a constant here: its gradient is produced by the Linear it belongs to (glue._FastLinearFn
    with ``add_bias=False``)."""

    @staticmethod
    def forward(ctx, x, r1, r2, gamma, beta, eps: float, relu: bool, owner, xbias=None, pos=None, copies: int = 0):
        C = x.shape[-1]
        xc = _f32c(x, "x")
        r1c = _f32c(r1, "residual") if r1 is not None else None
        r2c = _f32c(r2, "residual2") if r2 is not None else None
        for r in (r1c, r2c):
            if r is not None and r.shape != xc.shape:
                raise ValueError(f"add_layernorm: residual {tuple(r.shape)} vs x {tuple(xc.shape)}")
        rows = xc.numel() // C
        y = torch.empty_like(xc)
        posc = _f32c(pos, "pos") if pos is not None else None
        if posc is not None and posc.shape != xc.shape:
            raise ValueError(f"add_layernorm: pos {tuple(posc.shape)} vs x {tuple(xc.shape)}")
        y2 = torch.empty_like(xc) if posc is not None else None
        if not 0 <= copies <= 2:
            raise ValueError("add_layernorm: at most 2 extra copies of the output")
        ycs = [torch.empty_like(xc) for _ in range(copies)]
        need_s = r1c is not None or xbias is not None
        s = torch.empty_like(xc) if need_s else xc
        stats = torch.empty(2, rows, device=xc.device, dtype=torch.float32)
        st = _lib.load().gd4d_add_layernorm_fwd(
            xc.data_ptr(), _ptr(xbias), _ptr(r1c), _ptr(r2c), gamma.data_ptr(), beta.data_ptr(), _ptr(posc),
            y.data_ptr(), _ptr(y2), ycs[0].data_ptr() if copies > 0 else None, ycs[1].data_ptr() if copies > 1 else None,
            s.data_ptr() if need_s else None, stats[0].data_ptr(), stats[1].data_ptr(),
            rows, C, eps, int(relu), _stream_ptr(xc.device))
        _lib.check(st, "gd4d_add_layernorm_fwd")
        _count()
        ctx.save_for_backward(s, stats, gamma, beta)
        ctx.relu, ctx.owner = relu, owner
        ctx.has = (r1 is not None, r2 is not None)
        ctx.has_y2 = y2 is not None
        ctx.set_materialize_grads(False)                   # an unused output's gradient arrives as None
        outs = (y,) + ((y2,) if y2 is not None else ()) + tuple(ycs)
        return outs if len(outs) > 1 else y

    @staticmethod
    def backward(ctx, gy, *rest):
        s, stats, gamma, beta = ctx.saved_tensors
        C = s.shape[-1]
        rows = s.numel() // C
        rest = list(rest)
        gy2 = rest.pop(0) if ctx.has_y2 else None
        gcs = rest + [None] * (2 - len(rest))              # gradients of the copies of y
        g_pos = gy2                                        # d(y + pos)/dpos = 1
        if gy is None and gy2 is None and gcs[0] is None and gcs[1] is None:
            return (None,) * 11
        gy = _f32c(gy, "grad") if gy is not None else None
        gy2 = _f32c(gy2, "grad2") if gy2 is not None else None
        gcs = [_f32c(g, "grad copy") if g is not None else None for g in gcs]
        n_in = sum(g is not None for g in (gy, gy2, *gcs))
        gs = torch.empty_like(s)
        need_wgrad = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        gm = torch.empty_like(s) if ((ctx.relu or n_in > 1 or gy is None) and need_wgrad) else None
        st = _lib.load().gd4d_add_layernorm_bwd(
            _ptr(gy), _ptr(gy2), _ptr(gcs[0]), _ptr(gcs[1]), s.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
            gamma.data_ptr(),
            beta.data_ptr(), gs.data_ptr(), _ptr(gm), rows, C, int(ctx.relu), _stream_ptr(s.device))
        _lib.check(st, "gd4d_add_layernorm_bwd")
        _count()
        g_eff = gm if gm is not None else gy               # effective incoming gradient (masked / summed)
        dgamma = dbeta = None
        if need_wgrad:
            q = DeferredWgrad._active
            G, X = g_eff.reshape(rows, C), s.reshape(rows, C)
            mean, rstd = stats[0].reshape(rows, 1), stats[1].reshape(rows, 1)
            if q is not None and ctx.owner is not None:
                q.ln_items.append((*ctx.owner, G, X, mean, rstd))
            else:
                dgamma = (G * ((X - mean) * rstd)).sum(0)
                dbeta = G.sum(0)
        return (gs, gs if ctx.has[0] else None, gs if ctx.has[1] else None, dgamma, dbeta, None, None, None,
                None, g_pos, None)


def add_layernorm(x: torch.Tensor, ln: torch.nn.LayerNorm, r1: Optional[torch.Tensor] = None,
                  r2: Optional[torch.Tensor] = None, relu: bool = False,
                  xbias: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None, copies: int = 0):
    """[relu](ln(x + xbias + r1 + r2)) in one launch (forward) / one launch (backward dX).
    ``xbias``: the (C,) bias of the Linear that produced ``x`` with ``add_bias=False``.
    ``pos``: also return ``y + pos`` (the next attention block's query + query_pos) -> (y, y + pos).
    ``copies`` (0..2): also return that many identical copies of ``y`` (after ``y + pos`` if requested), ONE PER
    FURTHER CONSUMER of the result: every consumer then sends its own gradient back and the backward kernel sums
    them in registers; with a single output autograd would launch an elementwise add per extra consumer."""
    if r1 is None and r2 is not None:
        r1, r2 = r2, None
    owner = (ln.weight, ln.bias) if ln.weight.requires_grad else None
    if xbias is not None:
        xbias = _f32c(xbias.detach(), "xbias")
    return _AddLayerNormFn.apply(x, r1, r2, ln.weight, ln.bias, float(ln.eps), bool(relu), owner, xbias, pos, int(copies))


def can_fuse_layernorm(x: torch.Tensor, ln: torch.nn.LayerNorm) -> bool:
    C = x.shape[-1]
    return (ENABLED and x.is_cuda and x.dtype == torch.float32 and ln.weight is not None and ln.bias is not None
            and len(ln.normalized_shape) == 1 and C % 128 == 0 and C <= 1024
            and ln.weight.dtype == torch.float32)
