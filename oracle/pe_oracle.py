"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the frustum position-embedding input of
Graph-DETR4D's PE head:

    Detr3DHeadPE.position_embeding      dense_heads/detr3d_head_pe.py:427-491

up to, not including, the ``position_encoder`` 1x1 convolutions (library code): for every
level, camera, pixel and depth bin the pixel's frustum point is lifted through img2lidar,
normalised by pc_range, flagged when outside [0,1], laid out as (B*N, D*3, H, W) and passed
through inverse_sigmoid.  Op-by-op in fp32 torch-on-CPU in the reference's order, with the
4x4 mat-vec written out elementwise (sequential, no BLAS) like oracle/xview_oracle.py.

Pinned: tests/test_pe_oracle.py executes the unmodified reference method (oracle/ref_loader.
load_position_embeding) in the build container and compares; tests/golden/frustum_pe.npz
freezes its outputs for the GPU box.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this file.
"""
from __future__ import annotations

import numpy as np
import torch


def inverse_sigmoid(x, eps=1e-5):
    """mmdet 2.x models/utils/transformer.py::inverse_sigmoid == detr3d_transformer.py:28-43."""
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def img2lidar_fp32(img_metas):
    """np.linalg.inv of every float64 lidar2img, THEN the cast to fp32 (detr3d_head_pe.py:461-467)."""
    mats = [[np.linalg.inv(m) for m in meta["lidar2img"]] for meta in img_metas]
    return torch.as_tensor(np.asarray(mats)).to(torch.float32)            # (B,N,4,4)


def frustum_pe_input(level_shapes, img_metas, depth_num, depth_start, pc_range, masks=None):
    """Returns ([x_l (B*N, D*3, H_l, W_l) fp32], [mask_l (B,N,H_l,W_l) bool]) -- the tensors the
    reference feeds to ``self.position_encoder`` (:486) and returns as ``coords_masks`` (:489)."""
    eps = 1e-5
    pad_h, pad_w, _ = img_metas[0]["pad_shape"][0]                                  # :430
    i2l = img2lidar_fp32(img_metas)                                                 # :461-467
    B, N = i2l.shape[:2]
    xs, ms = [], []
    for lvl, (H, W) in enumerate(level_shapes):
        coords_h = torch.arange(H).float() * pad_h / H                             # :439
        coords_w = torch.arange(W).float() * pad_w / W                             # :440
        index = torch.arange(0, depth_num, 1).float()                              # :452
        index_1 = index + 1
        bin_size = (pc_range[3] - depth_start) / (depth_num * (1 + depth_num))     # :454 (python double)
        coords_d = depth_start + bin_size * index * index_1                        # :455
        D = depth_num
        cw = coords_w.view(W, 1, 1).expand(W, H, D)
        ch = coords_h.view(1, H, 1).expand(W, H, D)
        cd = coords_d.view(1, 1, D).expand(W, H, D)
        scale = torch.maximum(cd, torch.ones_like(cd) * eps)                       # :460
        px, py, pz, pw = cw * scale, ch * scale, cd, torch.ones_like(cd)
        M = i2l.view(B, N, 1, 1, 1, 16)
        out = []
        for r in range(3):                                                         # :468 matmul, rows 0..2 (:468 [..., :3])
            acc = M[..., 4 * r + 0] * px
            acc = acc + M[..., 4 * r + 1] * py
            acc = acc + M[..., 4 * r + 2] * pz
            acc = acc + M[..., 4 * r + 3] * pw
            out.append(acc)
        c3 = torch.stack(out, -1)                                                  # (B,N,W,H,D,3)
        for i in range(3):                                                         # :469-474
            c3[..., i] = (c3[..., i] - pc_range[i]) / (pc_range[3 + i] - pc_range[i])
        m = (c3 > 1.0) | (c3 < 0.0)                                                # :476
        m = m.flatten(-2).sum(-1) > (D * 0.5)                                      # :477
        m = m.permute(0, 1, 3, 2)                                                  # (B,N,H,W)
        if masks is not None:
            m = masks[lvl] | m                                                     # :478
        x = c3.permute(0, 1, 4, 5, 3, 2).contiguous().view(B * N, -1, H, W)        # :479
        xs.append(inverse_sigmoid(x))                                              # :480
        ms.append(m)
    return xs, ms
