"""Algorithmic-byte bookkeeping for the fused sampling kernels.

bytes_fwd = S * Ch * e + B*Q*C*4 + W_bytes + B*Q*3*4 + B*N*64
  S       = number of in-bounds bilinear corner reads over every valid
            (b, q, head, camera, level, point) sample, counted exactly here with
            torch ops on the device (it only feeds the roofline arithmetic; the
            1-ulp differences from the kernel's non-FMA projection move S by
            O(1e-6) relative)
  Ch * e  = bytes of one head slice (32 ch: fp32 128 B, bf16 64 B)
  W_bytes = raw weight / offset tensors read
bytes_bwd = B*Q*C*4 + S*Ch*e + 2*S*Ch*4 + grad outputs (same sizes as W_bytes)
            (the dense zero-fill of the grad map is a separate memset, counted
            by the caller when it is part of the timed region)
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

from .ops import MODE_C


@torch.no_grad()
def count_corner_reads(mode: int, shapes: Sequence[Tuple[int, int]], ref, offsets, lidar2img, pc_range,
                       img_h: float, img_w: float, num_heads: int, num_points: int) -> Dict[str, float]:
    B, Q = ref.shape[:2]
    N = lidar2img.shape[1]
    lo = ref.new_tensor(pc_range[:3])
    span = ref.new_tensor([pc_range[3] - pc_range[0], pc_range[4] - pc_range[1], pc_range[5] - pc_range[2]])
    pts = ref * span + lo                                          # (B,Q,3)
    if mode == MODE_C:
        pts = pts.view(B, Q, 1, 1, 3) + offsets.view(B, Q, num_heads, num_points, 3)
        heads = num_heads
    else:
        pts = pts.view(B, Q, 1, 1, 3)
        heads = 1                                                  # projection shared by all slices
    M = pts.shape[2] * pts.shape[3]
    pts = pts.reshape(B, 1, Q * M, 3)
    m = lidar2img.view(B, N, 1, 16)
    cam = [m[..., 4 * r + 0] * pts[..., 0] + m[..., 4 * r + 1] * pts[..., 1] + m[..., 4 * r + 2] * pts[..., 2]
           + m[..., 4 * r + 3] for r in range(3)]
    den = cam[2].clamp_min(1e-5)
    u = cam[0] / den / img_w
    v = cam[1] / den / img_h
    valid = (cam[2] > 1e-5) & (u > 0) & (u < 1) & (v > 0) & (v < 1)
    corners = 0
    unique_rows, per_level = 0, []
    img = (torch.arange(B, device=ref.device).view(B, 1, 1) * N + torch.arange(N, device=ref.device).view(1, N, 1))
    img = img.expand(B, N, u.shape[-1])
    for (H, W) in shapes:
        ix, iy = u * W - 0.5, v * H - 0.5
        x0, y0 = ix.floor(), iy.floor()
        cx = ((x0 >= 0) & (x0 <= W - 1)).int() + ((x0 + 1 >= 0) & (x0 + 1 <= W - 1)).int()
        cy = ((y0 >= 0) & (y0 <= H - 1)).int() + ((y0 + 1 >= 0) & (y0 + 1 <= H - 1)).int()
        n_l = int((cx * cy * valid).sum().item())
        corners = corners + n_l
        # distinct pixel rows touched in this level (the unique-footprint lower bound of SURVEY 8d)
        ids = []
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                ok = valid & (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
                ids.append(((img * H + yy.long().clamp(0, H - 1)) * W + xx.long().clamp(0, W - 1))[ok])
        u_l = int(torch.unique(torch.cat(ids)).numel())
        unique_rows += u_l
        per_level.append(dict(h=H, w=W, corner_reads=n_l, unique_rows=u_l, rows_in_level=B * N * H * W))
    n_valid = int(valid.sum().item())
    return dict(valid_samples=n_valid, valid_fraction=n_valid / valid.numel(),
                corner_reads_per_slice_group=int(corners), heads_counted=heads,
                unique_rows=int(unique_rows), per_level=per_level)


def algorithmic_bytes(mode: int, stats: Dict[str, float], B: int, Q: int, N: int, C: int, num_heads: int,
                      L: int, P: int, elem_bytes: int, wide: bool = False) -> Dict[str, float]:
    """Returns forward / backward algorithmic bytes for one launch."""
    slice_bytes = 32 * elem_bytes
    if mode == MODE_C and wide:
        # gather-then-project: every corner read is a whole C-channel row, the output is
        # (B,Hh,Q,C) (+ wsum), the feature gradient is a C-wide fp32 read-modify-write
        S = stats["corner_reads_per_slice_group"]
        row = C * elem_bytes
        w_bytes = B * Q * num_heads * (L * P + 3 * P) * 4 + B * Q * N * 4
        common = B * Q * 3 * 4 + B * N * 64
        out_b = B * Q * num_heads * (C + 1) * 4
        fwd = S * row + out_b + w_bytes + common
        bwd = out_b + S * row + 2 * S * C * 4 + 2 * w_bytes + common
        # SURVEY 8d AS WRITTEN (the reference op's bytes: one 32-channel head slice per corner, (B,Q,C) out)
        fwd_8d = S * slice_bytes + B * Q * C * 4 + w_bytes + common
        bwd_8d = B * Q * C * 4 + S * slice_bytes + 2 * S * 32 * 4 + 2 * w_bytes + common
        # unique-footprint lower bound: every distinct pixel row touched moves once (read; fp32 RMW in bwd)
        U = stats.get("unique_rows", 0)
        fwd_u = U * row + out_b + w_bytes + common
        bwd_u = out_b + U * row + 2 * U * C * 4 + 2 * w_bytes + common
        return dict(S=S, fwd=float(fwd), bwd=float(bwd), gather=float(S * row), red=float(S * C * 4),
                    fwd_8d=float(fwd_8d), bwd_8d=float(bwd_8d), fwd_unique=float(fwd_u), bwd_unique=float(bwd_u))
    if mode == MODE_C:
        S = stats["corner_reads_per_slice_group"]                  # already per (head, point)
        w_bytes = B * Q * num_heads * (L * P + 3 * P) * 4 + B * Q * N * 4
    else:
        S = stats["corner_reads_per_slice_group"] * (C // 32)      # every 32-ch slice reads each corner
        w_bytes = B * Q * N * L * P * 4
    common = B * Q * 3 * 4 + B * N * 64
    fwd = S * slice_bytes + B * Q * C * 4 + w_bytes + common
    bwd = B * Q * C * 4 + S * slice_bytes + 2 * S * 32 * 4 + 2 * w_bytes + common
    U = stats.get("unique_rows", 0) * (num_heads if mode == MODE_C else C // 32)   # slices per distinct row (upper bound)
    fwd_u = min(U, S) * slice_bytes + B * Q * C * 4 + w_bytes + common
    bwd_u = B * Q * C * 4 + min(U, S) * (slice_bytes + 2 * 32 * 4) + 2 * w_bytes + common
    return dict(S=S, fwd=float(fwd), bwd=float(bwd), gather=float(S * slice_bytes), red=float(S * 32 * 4),
                fwd_8d=float(fwd), bwd_8d=float(bwd), fwd_unique=float(fwd_u), bwd_unique=float(bwd_u))
