"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the feature position-embedding block of
Graph-DETR4D's PE head (second half):

    Detr3DHeadPE.forward              dense_heads/detr3d_head_pe.py:510-553
    SELayer.forward                   dense_heads/detr3d_head_pe.py:239-243
    SinePositionalEncoding3D.forward  models/utils/positional_encoding.py:58-100

Plain fp32 torch on the CPU, op by op in the reference's order.  The frustum part
(``position_embeding``, :427-491) is oracle/pe_oracle.py.  Pinned: tests/test_fpe_oracle.py executes
the reference's own lines (oracle/ref_loader.load_fpe_block) in the build container and compares;
tests/golden/fpe_block.npz freezes their outputs for the GPU box.  Only tests/ may import this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import pe_oracle


def level_masks(batch_size, num_cams, level_shapes, img_metas):
    """:519-536 -- ones everywhere, zeros over each camera's unpadded image, nearest-interpolated."""
    pad_h, pad_w, _ = img_metas[0]["pad_shape"][0]
    full = torch.ones((batch_size, num_cams, pad_h, pad_w))
    for b in range(batch_size):
        for n in range(num_cams):
            img_h, img_w, _ = img_metas[b]["img_shape"][n]
            full[b, n, :img_h, :img_w] = 0
    return [F.interpolate(full, size=tuple(s)).to(torch.bool) for s in level_shapes]


def sine_pe3d(mask, num_feats=128, temperature=10000, normalize=True, scale=2 * math.pi, eps=1e-6, offset=-0.5):
    """positional_encoding.py:58-100, mask (B,N,H,W) bool -> (B,N,3*num_feats,H,W)."""
    not_mask = 1 - mask.to(torch.int)
    n_embed = not_mask.cumsum(1, dtype=torch.float32)
    y_embed = not_mask.cumsum(2, dtype=torch.float32)
    x_embed = not_mask.cumsum(3, dtype=torch.float32)
    if normalize:
        n_embed = (n_embed + offset) / (n_embed[:, -1:, :, :] + eps) * scale
        y_embed = (y_embed + offset) / (y_embed[:, :, -1:, :] + eps) * scale
        x_embed = (x_embed + offset) / (x_embed[:, :, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    B, N, H, W = mask.size()
    outs = []
    for e in (n_embed, y_embed, x_embed):
        p = e[:, :, :, :, None] / dim_t
        outs.append(torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).view(B, N, H, W, -1))
    return torch.cat(outs, dim=4).permute(0, 1, 4, 2, 3)


def dim_t(num_feats=128, temperature=10000):
    d = torch.arange(num_feats, dtype=torch.float32)
    return temperature ** (2 * (d // 2) / num_feats)


def se_gate(sd, prefix, x, x_se):
    """SELayer (:239-243): x * sigmoid(conv_expand(relu(conv_reduce(x_se))))."""
    g = F.conv2d(x_se, sd[f"{prefix}.conv_reduce.weight"], sd[f"{prefix}.conv_reduce.bias"])
    g = F.conv2d(F.relu(g), sd[f"{prefix}.conv_expand.weight"], sd[f"{prefix}.conv_expand.bias"])
    return x * g.sigmoid()


def _seq(sd, prefix, x):
    x = F.relu(F.conv2d(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"]))
    return F.conv2d(x, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"])


def fpe_block(sd, mlvl_feats, img_metas, depth_num, depth_start, pc_range, with_detach=True, num_feats=128):
    """:510-553 -> (new mlvl_feats list, masks list).  ``sd`` = state_dict of a head holding
    position_encoder / adapt_pos3d / fpe (the names the reference uses)."""
    feats = list(mlvl_feats)
    if with_detach:                                                        # :512-516, level 0 only, 6 current cams
        feats[0] = torch.cat([feats[0][:, :6], feats[0][:, 6:].detach()], 1)
    B, N = feats[0].shape[:2]
    shapes = [tuple(f.shape[-2:]) for f in feats]
    masks = level_masks(B, N, shapes, img_metas)
    xs, _ = pe_oracle.frustum_pe_input(shapes, img_metas, depth_num, depth_start, pc_range, masks)
    out = []
    for l, f in enumerate(feats):
        pe = _seq(sd, "position_encoder", xs[l])                           # :486
        pe = se_gate(sd, "fpe", pe, f.flatten(0, 1)).view(f.size())       # :545
        sin = sine_pe3d(masks[l], num_feats=num_feats)                     # :550
        sin = _seq(sd, "adapt_pos3d", sin.flatten(0, 1)).view(f.size())    # :551
        out.append(f + (pe + sin))                                         # :552-553
    return out, masks
