"""Host side of the fused frustum position-embedding input (include/gd4d_frustum.h;
SURVEY.md 8f row f4): the elementwise body of

    Detr3DHeadPE.position_embeding   projects/mmdet3d_plugin/models/dense_heads/detr3d_head_pe.py:427-491

which sits immediately upstream of the cross-view sampling path in every Graph-DETR4D config.
``position_embeding`` keeps the reference's signature and return value (the position_encoder
convolutions are library code and are passed in); ``frustum_position_input`` returns the tensors
those convolutions consume.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ops import _count, _stream_ptr


def img2lidar_to_tensor(img_metas, device) -> torch.Tensor:
    """float32(np.linalg.inv(float64 lidar2img)) per camera, exactly as the reference builds it
    (detr3d_head_pe.py:461-467: inverse in numpy's dtype first, the fp32 cast second) -> (B,N,4,4)."""
    mats = np.asarray([[np.linalg.inv(m) for m in meta["lidar2img"]] for meta in img_metas])
    return torch.as_tensor(mats.astype(np.float32)).to(device, non_blocking=True)


def frustum_position_input(level_shapes: Sequence[Tuple[int, int]], img_metas, depth_num: int, depth_start,
                           pc_range, masks: Optional[Sequence[torch.Tensor]] = None, device="cuda",
                           img2lidar: Optional[torch.Tensor] = None):
    """-> ([x_l (B*N, 3*depth_num, H_l, W_l) fp32], [mask_l (B,N,H_l,W_l) bool]): the input of
    ``position_encoder`` (:486) and ``coords_masks`` (:489), ONE launch for all levels."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("frustum_position_input runs on CUDA only (the CPU oracle lives in oracle/, test-only)")
    i2l = img2lidar if img2lidar is not None else img2lidar_to_tensor(img_metas, device)
    if i2l.dtype != torch.float32 or not i2l.is_cuda or i2l.dim() != 4 or tuple(i2l.shape[2:]) != (4, 4):
        raise ValueError("img2lidar must be a CUDA float32 (B,N,4,4) tensor")
    i2l = i2l.contiguous()
    B, N = int(i2l.shape[0]), int(i2l.shape[1])
    pad_h, pad_w = img_metas[0]["pad_shape"][0][0:2]                          # :430, sample 0 / cam 0 for all
    D = int(depth_num)
    bin_size = (pc_range[3] - depth_start) / (D * (1 + D))                    # python double (:454), fp32 at the call
    lo_span = (C.c_float * 6)(*[float(pc_range[i]) for i in range(3)],
                              *[float(pc_range[3 + i] - pc_range[i]) for i in range(3)])
    lib = _lib.load()
    xs: List[torch.Tensor] = []
    ms: List[torch.Tensor] = []
    m_ins = []
    for lvl, (H, W) in enumerate(level_shapes):
        H, W = int(H), int(W)
        m_in = None
        if masks is not None:
            m_in = masks[lvl]
            if tuple(m_in.shape) != (B, N, H, W) or not m_in.is_cuda:
                raise ValueError(f"masks[{lvl}] must be a CUDA (B,N,H,W)=({B},{N},{H},{W}) tensor")
            m_in = m_in.to(torch.uint8).contiguous()
        m_ins.append(m_in)
        xs.append(torch.empty((B * N, 3 * D, H, W), device=device, dtype=torch.float32))
        ms.append(torch.empty((B, N, H, W), device=device, dtype=torch.uint8))
    # every level in ONE launch (the coarse levels alone are a few dozen CTAs each); chunks of 8 levels
    for l0 in range(0, len(xs), 8):
        n = min(8, len(xs) - l0)
        ptr = lambda ts: (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts[l0:l0 + n]])
        hs = (C.c_int32 * n)(*[int(h) for h, _ in level_shapes[l0:l0 + n]])
        ws = (C.c_int32 * n)(*[int(w) for _, w in level_shapes[l0:l0 + n]])
        st = lib.gd4d_frustum_pe_levels(i2l.data_ptr(), ptr(m_ins), ptr(xs), ptr(ms), B * N, n, hs, ws, D,
                                        float(pad_h), float(pad_w), float(depth_start), float(bin_size), lo_span,
                                        _stream_ptr(device))
        _lib.check(st, "gd4d_frustum_pe_levels")
        _count()
    ms = [m.view(torch.bool) for m in ms]
    return xs, ms


def position_embeding(img_feats: Sequence[torch.Tensor], img_metas, masks=None, *, position_encoder,
                      depth_num: int = 64, depth_start=1, pc_range=None, embed_dims: Optional[int] = None):
    """Drop-in for ``Detr3DHeadPE.position_embeding(img_feats, img_metas, masks)``: returns
    (coords_position_embedings [(B,N,embed_dims,H,W)], coords_masks [(B,N,H,W) bool])."""
    B, N = int(img_feats[0].shape[0]), int(img_feats[0].shape[1])
    shapes = [(int(f.shape[-2]), int(f.shape[-1])) for f in img_feats]
    xs, ms = frustum_position_input(shapes, img_metas, depth_num, depth_start, pc_range, masks,
                                    device=img_feats[0].device)
    embs = []
    for x, (H, W) in zip(xs, shapes):
        e = position_encoder(x)                                                # :486 (library 1x1 convs)
        embs.append(e.view(B, N, embed_dims if embed_dims is not None else e.shape[1], H, W))
    return embs, ms
