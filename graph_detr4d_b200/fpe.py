"""Host side of the feature position-embedding block of the PE head (include/gd4d_fpe.h; SURVEY.md 8f
row f4, second half) -- a drop-in for lines 510-553 of

    Detr3DHeadPE.forward   projects/mmdet3d_plugin/models/dense_heads/detr3d_head_pe.py

    mlvl_feats = position_embed_features(self, mlvl_feats, img_metas)

``self`` is the head: it must carry the sub-modules the reference builds (``position_encoder``,
``adapt_pos3d``, ``fpe`` with ``conv_reduce`` / ``conv_expand``, ``positional_encoding`` with its
``num_feats / temperature / normalize / scale / eps / offset``) and ``depth_num``, ``depth_start``,
``pc_range``, ``with_detach``.  The 1x1 convolutions stay library calls; everything elementwise around
them is ours:

  * level padding masks straight from ``img_shape`` (one launch per level; the reference fills a
    full-resolution (B,N,pad_h,pad_w) mask in a python double loop and interpolates it)
  * frustum position-embedding input (frustum.py, one launch per level)
  * 3-D sine embedding from the image sizes (one launch per level instead of 3 cumsums + ~20 ops)
  * ``feat + (pe * sigmoid(gate) + sine)`` as one launch forward, one backward
  * the level-0 past-frame detach (:512-516) as an aliasing autograd node: no concatenation copy of
    the largest level, the gradient of cameras 6.. is zeroed on the way back instead.
CUDA only, no fallback.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from .frustum import frustum_position_input
from .ops import _count, _stream_ptr

__all__ = ["level_masks", "sine_pe3d", "fpe_combine", "detach_past_frames", "position_embed_features"]


def _img_hw_tensor(img_metas, num_cams: int, device) -> torch.Tensor:
    hw = np.asarray([[meta["img_shape"][n][0:2] for n in range(num_cams)] for meta in img_metas], dtype=np.int32)
    return torch.from_numpy(hw).to(device, non_blocking=True)            # (B,N,2)


def _require_cuda(device):
    if torch.device(device).type != "cuda":
        raise RuntimeError("the position-embedding block runs on CUDA only (the CPU oracle lives in oracle/, test-only)")


def level_masks(level_shapes: Sequence[Sequence[int]], img_metas, num_cams: int, device="cuda",
                img_hw: torch.Tensor = None) -> List[torch.Tensor]:
    """detr3d_head_pe.py:519-536 -> [(B,N,H_l,W_l) bool], True = padding."""
    _require_cuda(device)
    hw = img_hw if img_hw is not None else _img_hw_tensor(img_metas, num_cams, device)
    B, N = int(hw.shape[0]), int(hw.shape[1])
    pad_h, pad_w = (int(v) for v in img_metas[0]["pad_shape"][0][0:2])
    lib, out = _lib.load(), []
    for (H, W) in level_shapes:
        m = torch.empty((B, N, int(H), int(W)), device=hw.device, dtype=torch.uint8)
        _lib.check(lib.gd4d_level_mask(hw.data_ptr(), m.data_ptr(), B * N, int(H), int(W), pad_h, pad_w,
                                       _stream_ptr(hw.device)), "gd4d_level_mask")
        _count()
        out.append(m.view(torch.bool))
    return out


_DIM_T = {}


def _dim_t(num_feats: int, temperature, device) -> torch.Tensor:
    key = (int(num_feats), float(temperature), torch.device(device))
    t = _DIM_T.get(key)
    if t is None:
        d = torch.arange(num_feats, dtype=torch.float32)                   # positional_encoding.py:82-84, on the host
        t = (temperature ** (2 * (d // 2) / num_feats)).to(device)
        _DIM_T[key] = t
    return t


def sine_pe3d(level_shape, img_metas, num_cams: int, num_feats: int = 128, temperature=10000, normalize=True,
              scale=2 * np.pi, eps=1e-6, offset=0.0, device="cuda", img_hw: torch.Tensor = None) -> torch.Tensor:
    """SinePositionalEncoding3D(mask of this level) -> (B*N, 3*num_feats, H, W) fp32, without
    materialising the mask or its cumsums (positional_encoding.py:58-100)."""
    _require_cuda(device)
    hw = img_hw if img_hw is not None else _img_hw_tensor(img_metas, num_cams, device)
    B, N = int(hw.shape[0]), int(hw.shape[1])
    H, W = int(level_shape[0]), int(level_shape[1])
    pad_h, pad_w = (int(v) for v in img_metas[0]["pad_shape"][0][0:2])
    out = torch.empty((B * N, 3 * num_feats, H, W), device=hw.device, dtype=torch.float32)
    dt = _dim_t(num_feats, temperature, hw.device)
    _lib.check(_lib.load().gd4d_sine_pe3d(hw.data_ptr(), dt.data_ptr(), out.data_ptr(), B, N, H, W, pad_h, pad_w,
                                          int(num_feats), 1 if normalize else 0, float(scale), float(eps),
                                          float(offset), _stream_ptr(hw.device)), "gd4d_sine_pe3d")
    _count()
    return out


class _CombineFn(torch.autograd.Function):
    """feat + (pe * sigmoid(gate) + sine): one launch forward, one backward (SELayer :243, :552-553)."""

    @staticmethod
    def forward(ctx, feat, pe, gate, sine):
        feat, pe, gate, sine = (t.contiguous() for t in (feat, pe, gate, sine))
        out = torch.empty_like(feat)
        _lib.check(_lib.load().gd4d_fpe_combine_fwd(feat.data_ptr(), pe.data_ptr(), gate.data_ptr(), sine.data_ptr(),
                                                    out.data_ptr(), feat.numel(), _stream_ptr(feat.device)),
                   "gd4d_fpe_combine_fwd")
        _count()
        ctx.save_for_backward(pe, gate)
        return out

    @staticmethod
    def backward(ctx, g):
        pe, gate = ctx.saved_tensors
        g = g.contiguous()
        need_pe, need_gate = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gpe = torch.empty_like(pe) if need_pe else None
        ggate = torch.empty_like(gate) if need_gate else None
        if need_pe or need_gate:
            _lib.check(_lib.load().gd4d_fpe_combine_bwd(g.data_ptr(), pe.data_ptr(), gate.data_ptr(),
                                                        gpe.data_ptr() if need_pe else None,
                                                        ggate.data_ptr() if need_gate else None, g.numel(),
                                                        _stream_ptr(g.device)), "gd4d_fpe_combine_bwd")
            _count()
        return (g if ctx.needs_input_grad[0] else None, gpe, ggate, g if ctx.needs_input_grad[3] else None)


def fpe_combine(feat, pe, gate, sine):
    for t in (feat, pe, gate, sine):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("fpe_combine needs CUDA float32 tensors (no CPU fallback)")
        if t.shape != feat.shape:
            raise ValueError("feat, pe, gate and sine must have the same shape")
    return _CombineFn.apply(feat, pe, gate, sine)


class _DetachCamsFn(torch.autograd.Function):
    """torch.cat([x[:, :keep], x[:, keep:].detach()], 1) without the copy: forward aliases x, backward
    zeroes the gradient of cameras keep.. (detr3d_head_pe.py:512-516)."""

    @staticmethod
    def forward(ctx, x, keep: int):
        ctx.keep = keep
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        g[:, ctx.keep:].zero_()
        return g, None


def detach_past_frames(level0: torch.Tensor, current_cams: int = 6) -> torch.Tensor:
    if level0.shape[1] <= current_cams or not level0.requires_grad:
        return level0
    return _DetachCamsFn.apply(level0, current_cams)


def position_embed_features(head, mlvl_feats: Sequence[torch.Tensor], img_metas) -> List[torch.Tensor]:
    """Lines 510-553 of ``Detr3DHeadPE.forward``: returns the new ``mlvl_feats`` list."""
    feats = list(mlvl_feats)
    if not feats[0].is_cuda:
        raise RuntimeError("position_embed_features runs on CUDA tensors only (no CPU fallback)")
    if getattr(head, "with_detach", False):
        feats[0] = detach_past_frames(feats[0], 6)                             # :512-516
    B, N = int(feats[0].shape[0]), int(feats[0].shape[1])
    dev = feats[0].device
    shapes = [(int(f.shape[-2]), int(f.shape[-1])) for f in feats]
    hw = _img_hw_tensor(img_metas, N, dev)
    masks = level_masks(shapes, img_metas, N, dev, img_hw=hw)                  # :519-536
    xs, _ = frustum_position_input(shapes, img_metas, head.depth_num, head.depth_start, head.pc_range, masks,
                                   device=dev)                                 # :427-485
    pos = head.positional_encoding
    out = []
    for l, f in enumerate(feats):
        pe = head.position_encoder(xs[l])                                      # :486   library 1x1 convs
        flat = f.flatten(0, 1)
        gate = head.fpe.conv_expand(head.fpe.act1(head.fpe.conv_reduce(flat)))  # :240-242
        sine = sine_pe3d(shapes[l], img_metas, N, pos.num_feats, pos.temperature, pos.normalize, pos.scale,
                         pos.eps, pos.offset, dev, img_hw=hw)                  # :550
        sine = head.adapt_pos3d(sine)                                          # :551
        out.append(fpe_combine(flat.float(), pe, gate, sine).view(f.size()))   # :243, :552-553
    return out
