"""TEST / BASELINE INFRASTRUCTURE ONLY -- nn.Module wrappers around the CPU oracle
so the reference arm of bench.py (``--impl reference``) and the parity tests can
drop the oracle's port of the attention into the same decoder harness.  Parameter
names equal the reference's (detr3d_transformer.py:292-304,
deform3d_cross_attn.py:100-121).  Never imported by the product package."""
import math

import torch
import torch.nn as nn

from . import xview_oracle as xo


def _pos_enc(in_dims, c):
    return nn.Sequential(nn.Linear(in_dims, c), nn.LayerNorm(c), nn.ReLU(inplace=True),
                         nn.Linear(c, c), nn.LayerNorm(c), nn.ReLU(inplace=True))


class OracleDetr3DCrossAtten(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=5, num_cams=6,
                 im2col_step=64, pc_range=None, dropout=0.1, norm_cfg=None, init_cfg=None,
                 batch_first=False, **_):
        super().__init__()
        self.pc_range, self.num_heads = pc_range, num_heads
        self.attention_weights = nn.Linear(embed_dims, num_cams * num_levels * num_points)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.position_encoder = _pos_enc(3, embed_dims)
        nn.init.zeros_(self.attention_weights.weight); nn.init.zeros_(self.attention_weights.bias)
        nn.init.xavier_uniform_(self.output_proj.weight); nn.init.zeros_(self.output_proj.bias)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        sd = dict(self.named_parameters())
        return xo.detr3d_cross_atten_forward(sd, query, value, query_pos, reference_points,
                                             kwargs["img_metas"], self.pc_range)


class OracleDeform3DCrossAttn(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=5, num_cams=6,
                 im2col_step=64, pc_range=None, dropout=0.1, norm_cfg=None, init_cfg=None,
                 batch_first=False, fix_offset=False, depth_encode=False, **_):
        super().__init__()
        self.pc_range, self.num_heads, self.depth_encode = pc_range, num_heads, depth_encode
        self.cam_attention_weights = nn.Linear(embed_dims, num_cams)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.position_encoder = _pos_enc(4 if depth_encode else 3, embed_dims)
        self.deform_sampling_offsets = nn.Linear(embed_dims, num_heads * num_points * 3)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        for lin in (self.cam_attention_weights, self.attention_weights, self.deform_sampling_offsets):
            nn.init.zeros_(lin.weight); nn.init.zeros_(lin.bias)
        thetas = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin(), thetas.cos()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 3).repeat(1, 1, num_points, 1)
        for i in range(num_points):
            grid[:, :, i, :] *= i + 1
        self.deform_sampling_offsets.bias.data = grid.view(-1)
        for lin in (self.output_proj, self.value_proj):
            nn.init.xavier_uniform_(lin.weight); nn.init.zeros_(lin.bias)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        sd = dict(self.named_parameters())
        return xo.deform3d_cross_attn_forward(sd, query, value, query_pos, reference_points,
                                              kwargs["img_metas"], self.pc_range, self.num_heads,
                                              self.depth_encode)


def build_oracle_attention(cfg):
    cfg = dict(cfg)
    cls = {"Detr3DCrossAtten": OracleDetr3DCrossAtten, "Deform3DCrossAttn": OracleDeform3DCrossAttn}[cfg.pop("type")]
    return cls(**cfg)
