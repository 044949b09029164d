"""TEST INFRASTRUCTURE ONLY -- the reference-side decoder used to pin row a9:

  * ``Detr3DTransformerDecoder`` (projects/mmdet3d_plugin/models/utils/detr3d_transformer.py:151-225) is the
    reference's OWN class, loaded unmodified through oracle/ref_loader.py (its mmcv parent
    ``TransformerLayerSequence`` is the loader's shim: an nn.Module with an empty ``layers`` list) -- the
    layer loop, the logit-space reference-point refinement and the detach (:192-214) are EXECUTED, not restated;
  * its cross attention is the reference's own ``Detr3DCrossAtten`` / ``Deform3DCrossAttn`` class (the latter
    with the one documented token fix of its dead CPU branch, ref_loader._patch_dc);
  * the layer around them is mmcv 1.x ``DetrTransformerDecoderLayer`` / ``BaseTransformerLayer`` with
    operation_order ('self_attn','norm','cross_attn','norm','ffn','norm') (configs/detr3d/detr3d_res50.py:65-83)
    -- third-party, un-vendored, un-pinned (mmcv-full 1.x): its published algorithm is restated in
    ``MMCVDecoderLayer`` below (post-norm; ``MultiheadAttention`` wrapper = nn.MultiheadAttention on
    q = k = query + query_pos, v = query, + identity; ``FFN`` = Linear-ReLU-Dropout-Linear-Dropout + identity),
    with mmcv's sub-module names so state dicts are interchangeable with graph_detr4d_b200.decoder.

Only tests/ (and tests/golden/make_golden_decoder.py) may import this file.
"""
import torch
import torch.nn as nn


class MMCVMultiheadAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention (batch_first=False)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0, dropout_layer_p=0.0):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(dropout_layer_p)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        out = self.attn(query=query, key=key, value=value)[0]
        return identity + self.dropout_layer(self.proj_drop(out))


class MMCVFFN(nn.Module):
    """mmcv.cnn.bricks.transformer.FFN (num_fcs=2, ReLU, add_identity)."""

    def __init__(self, embed_dims, feedforward_channels, ffn_drop=0.0):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))
        self.dropout_layer = nn.Identity()

    def forward(self, x, identity=None):
        out = self.layers(x)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


class MMCVDecoderLayer(nn.Module):
    """mmcv BaseTransformerLayer, operation_order ('self_attn','norm','cross_attn','norm','ffn','norm')."""
    ORDER = ("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")

    def __init__(self, cross_attn, embed_dims=256, num_heads=8, feedforward_channels=512, dropout=0.0):
        super().__init__()
        self.attentions = nn.ModuleList([MMCVMultiheadAttention(embed_dims, num_heads, dropout, 0.0, dropout), cross_attn])
        self.ffns = nn.ModuleList([MMCVFFN(embed_dims, feedforward_channels, dropout)])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dims) for _ in range(3)])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, **kwargs):
        norm_i = 0
        for op in self.ORDER:
            if op == "self_attn":
                query = self.attentions[0](query, query, query, None, query_pos=query_pos, key_pos=query_pos)
            elif op == "norm":
                query = self.norms[norm_i](query)
                norm_i += 1
            elif op == "cross_attn":
                query = self.attentions[1](query, key, value, None, query_pos=query_pos, key_pos=key_pos, **kwargs)
            else:
                query = self.ffns[0](query, None)
        return query


def build_reference_decoder(ref, variant, num_cams, num_layers, embed_dims=64, num_heads=2, num_points=4,
                            feedforward_channels=128, pc_range=None):
    """``ref`` = oracle.ref_loader.load().  The reference's decoder class around reference attention classes."""
    dec = ref.Detr3DTransformerDecoder(None, num_layers, return_intermediate=True)
    for _ in range(num_layers):
        if variant == "A":
            attn = ref.Detr3DCrossAtten(embed_dims=embed_dims, num_heads=num_heads, num_levels=4, num_points=1,
                                        num_cams=num_cams, pc_range=pc_range, dropout=0.0)
        else:
            attn = ref.Deform3DCrossAttnCPU(embed_dims=embed_dims, num_heads=num_heads, num_levels=4,
                                            num_points=num_points, num_cams=num_cams, pc_range=pc_range, dropout=0.0)
        dec.layers.append(MMCVDecoderLayer(attn, embed_dims, num_heads, feedforward_channels, 0.0))
    return dec


def make_reg_branches(num_layers, embed_dims, code_size=10, num_reg_fcs=2):
    """detr3d_head.py:72-95."""
    branches = []
    for _ in range(num_layers):
        fcs = []
        for _ in range(num_reg_fcs):
            fcs += [nn.Linear(embed_dims, embed_dims), nn.ReLU()]
        fcs.append(nn.Linear(embed_dims, code_size))
        branches.append(nn.Sequential(*fcs))
    return nn.ModuleList(branches)
