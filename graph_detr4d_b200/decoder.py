"""Caller of the hot path: a thin re-statement of the reference's decoder loop
(SURVEY.md section 8a row a9), needed to run BASELINE.json configs 2-3.

  Detr3DTransformerDecoder.forward   detr3d_transformer.py:166-225
      6 x layer(query, key=None, value=mlvl_feats, query_pos, reference_points, img_metas)
      then ref[..., :2] += reg[..., :2]; ref[..., 2] += reg[..., 4] in logit space,
      sigmoid, DETACH (:201-214)
  Detr3DTransformer.forward          detr3d_transformer.py:128-147
      split query_embed -> (query_pos, query); ref = sigmoid(Linear(query_pos))

The layer itself is mmcv's DetrTransformerDecoderLayer with
operation_order ('self_attn','norm','cross_attn','norm','ffn','norm')
(projects/configs/detr3d/detr3d_res50.py:65-83): post-norm, nn.MultiheadAttention
with q = k = query + query_pos, FFN 256->512->256.  Sub-module names follow mmcv's
(`attentions.{0,1}`, `ffns.0.layers`, `norms.{0,1,2}`) so decoder checkpoints map
1:1.  Everything except the cross-attention is stock torch (cuBLAS / library) --
plumbing around the hot path, deliberately not re-implemented.

``cross_attn_factory`` lets the CPU baseline (bench.py --impl reference) build the
very same decoder around the oracle's port of the attention.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn

from . import fused
from .glue import can_defer_bias, fast_layer_norm, fast_linear, linear_relu, self_attention
from .modules import build_attention, inverse_sigmoid


# In-place refresh of the packed generator weights (one multi-tensor copy per forward instead of two
# concatenations per layer).  The packed copy is what the generator GEMM saves for its backward, so an
# in-place refresh is only safe when every forward is followed by its backward before the next forward
# -- which the CUDA-graph step (graphed.GraphedTrainStep) guarantees and sets this flag for.  Eager
# callers (two forwards before one backward: teacher/student, multi-sample losses) get a fresh
# concatenation per forward, like the reference's three separate Linears.
STATIC_GENERATOR_PACKS = False


def _dropout_active(m: nn.Module) -> bool:
    return isinstance(m, nn.Dropout) and m.training and m.p > 0


def _run_branch(seq: nn.Sequential, x):
    """Linear(-ReLU)* chain of a regression branch; Linear+ReLU pairs run as GEMM + one launch."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                x = linear_relu(x, m)
                i += 1
            else:
                x = fast_linear(x, m)
        else:
            x = m(x)
        i += 1
    return x


class SelfAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention, batch_first=False."""

    def __init__(self, embed_dims=256, num_heads=8, dropout=0.1):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout)
        self.dropout_layer = nn.Dropout(dropout)

    def branch(self, query, query_pos=None, defer_bias=False, qk_in=None):
        """dropout(attention(query)) without the identity.  ``defer_bias``: returns
        ``(out_without_out_proj_bias, bias)`` for a consumer that adds the bias itself (only when
        dropout is inactive, else ``(out, None)``)."""
        if query.is_cuda:
            defer = defer_bias and not _dropout_active(self.dropout_layer) and \
                can_defer_bias(query, self.attn.out_proj)
            out = self_attention(query, query_pos, self.attn, out_bias=not defer, qk_in=qk_in)   # glue.py
            if defer_bias:
                return (out, self.attn.out_proj.bias) if defer else (self.dropout_layer(out), None)
        else:
            q = k = query if query_pos is None else query + query_pos
            out = self.attn(q, k, value=query, need_weights=False)[0]
            if defer_bias:
                return self.dropout_layer(out), None
        return self.dropout_layer(out)

    def forward(self, query, query_pos=None):
        return query + self.branch(query, query_pos)


class FFN(nn.Module):
    """mmcv FFN: Linear-ReLU-Dropout-Linear-Dropout + identity."""

    def __init__(self, embed_dims=256, feedforward_channels=512, ffn_drop=0.1):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))

    def branch(self, x, defer_bias=False):
        h = self.layers[0][2](linear_relu(x, self.layers[0][0]))
        lin2 = self.layers[1]
        if defer_bias:
            if not _dropout_active(self.layers[2]) and can_defer_bias(h, lin2):
                return fast_linear(h, lin2, add_bias=False), lin2.bias
            return self.layers[2](fast_linear(h, lin2)), None
        return self.layers[2](fast_linear(h, lin2))

    def forward(self, x):
        return x + self.branch(x)


class DecoderLayer(nn.Module):
    def __init__(self, cross_attn: nn.Module, embed_dims=256, num_heads=8, feedforward_channels=512,
                 dropout=0.1):
        super().__init__()
        self.attentions = nn.ModuleList([SelfAttention(embed_dims, num_heads, dropout), cross_attn])
        self.ffns = nn.ModuleList([FFN(embed_dims, feedforward_channels, dropout)])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dims) for _ in range(3)])

    def forward(self, query, value, query_pos, reference_points, img_metas, q_plus_pos=None, want_next=False,
                query_res=None):
        """``q_plus_pos``: ``query + query_pos`` if the previous layer already produced it;
        ``want_next``: also return what the NEXT layer consumes -> (output, output + pos, output copy for the next
        layer's value projection, output copy for its residual); ``query_res``: the copy of ``query`` to use as this
        layer's first residual (each consumer of a fused LayerNorm's result gets its own copy, fused.add_layernorm)."""
        # post-norm layer: every "branch + identity" sum is folded into the LayerNorm that
        # follows it (fused.add_layernorm: one launch) when the tensors are CUDA fp32
        fuse = all(fused.can_fuse_layernorm(query, n) for n in self.norms)
        cross = self.attentions[1]
        if not (fuse and hasattr(cross, "forward_parts")):
            query = fast_layer_norm(self.attentions[0](query, query_pos), self.norms[0])
            query = cross(query, None, value, None, query_pos=query_pos,
                          reference_points=reference_points, img_metas=img_metas)
            query = fast_layer_norm(query, self.norms[1])
            out = fast_layer_norm(self.ffns[0](query), self.norms[2])
            return (out, None, out, out) if want_next else out
        # every LayerNorm whose output feeds an attention block also emits output + query_pos
        out, bias = self.attentions[0].branch(query, query_pos, defer_bias=True, qk_in=q_plus_pos)
        query, qp = fused.add_layernorm(out, self.norms[0], query if query_res is None else query_res, xbias=bias,
                                        pos=query_pos)
        out, res, pos, bias = cross.forward_parts(query, None, value, None, query_pos=query_pos,
                                                  reference_points=reference_points, img_metas=img_metas,
                                                  defer_bias=True, query_with_pos=qp)
        # two consumers (the FFN's first Linear, the next residual): one copy each, their gradients meet again
        # inside the LayerNorm backward kernel instead of in an elementwise add
        if not fused.LN_COPIES:                            # one output, autograd adds the consumers' gradients
            query = fused.add_layernorm(out, self.norms[1], res, pos, xbias=bias)
            out, bias = self.ffns[0].branch(query, defer_bias=True)
            if want_next:
                o, qp_next = fused.add_layernorm(out, self.norms[2], query, xbias=bias, pos=query_pos)
                return o, qp_next, o, o
            return fused.add_layernorm(out, self.norms[2], query, xbias=bias)
        query, query_r = fused.add_layernorm(out, self.norms[1], res, pos, xbias=bias, copies=1)
        out, bias = self.ffns[0].branch(query, defer_bias=True)
        if want_next:   # consumers: the stacked output, the next layer's value projection and its first residual
            o, qp_next, o_v, o_r = fused.add_layernorm(out, self.norms[2], query_r, xbias=bias, pos=query_pos, copies=2)
            return o, qp_next, o_v, o_r
        return fused.add_layernorm(out, self.norms[2], query_r, xbias=bias)


class Detr3DTransformerDecoder(nn.Module):
    def __init__(self, attn_cfg: dict, num_layers=6, embed_dims=256, num_heads=8,
                 feedforward_channels=512, dropout=0.1, return_intermediate=True,
                 cross_attn_factory: Optional[Callable[[dict], nn.Module]] = None):
        super().__init__()
        factory = cross_attn_factory or build_attention
        self.layers = nn.ModuleList([
            DecoderLayer(factory(dict(attn_cfg)), embed_dims, num_heads, feedforward_channels, dropout)
            for _ in range(num_layers)])
        self.return_intermediate = return_intermediate
        self.embed_dims = embed_dims

    @torch.no_grad()
    def _refresh_generator_packs(self, query):
        """One multi-tensor copy refreshes EVERY layer's packed generator weights (the three
        generator Linears run as one GEMM, modules.Deform3DCrossAttn) instead of two
        concatenations per layer."""
        if not (fused.ENABLED and query.is_cuda and query.dtype == torch.float32):
            return
        if not (STATIC_GENERATOR_PACKS or torch.cuda.is_current_stream_capturing()):
            return
        dsts, srcs, mods = [], [], []
        for layer in self.layers:
            m = layer.attentions[1]
            if hasattr(m, "generator_pack_slots"):
                d, s = m.generator_pack_slots()
                dsts += d
                srcs += [p.detach() for p in s]
                mods.append(m)
        if dsts:
            torch._foreach_copy_(dsts, srcs)
            for m in mods:
                m._gen_pack_fresh = True

    def forward(self, query, value, query_pos, reference_points, reg_branches=None, img_metas=None):
        output = query
        intermediate, intermediate_ref = [], []
        self._refresh_generator_packs(query)
        q_plus_pos = None
        out_v = out_r = output                     # what the next layer's value projection / first residual consume
        for lid, layer in enumerate(self.layers):
            if lid + 1 < len(self.layers):
                output, q_plus_pos, out_v, out_r = layer(out_v, value, query_pos, reference_points, img_metas,
                                                         q_plus_pos=q_plus_pos, want_next=True, query_res=out_r)
            else:
                output = layer(out_v, value, query_pos, reference_points, img_metas, q_plus_pos=q_plus_pos,
                               query_res=out_r)
            if reg_branches is not None:                                    # :201-214
                tmp = output.permute(1, 0, 2)
                tmp = _run_branch(reg_branches[lid], tmp)
                if fused.ENABLED and tmp.is_cuda and tmp.dtype == torch.float32:
                    reference_points = fused.ref_update(tmp, reference_points)      # one launch
                else:
                    new_ref = torch.zeros_like(reference_points)
                    new_ref[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points[..., :2])
                    new_ref[..., 2:3] = tmp[..., 4:5] + inverse_sigmoid(reference_points[..., 2:3])
                    reference_points = new_ref.sigmoid().detach()
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_ref.append(reference_points)
        # the per-forward pack of the feature maps has served all layers: drop the cache's strong
        # references (autograd keeps what backward needs), so one step's maps do not outlive it
        from . import modules as _modules
        _modules.clear_pack_cache()
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_ref)
        return output, reference_points


class Detr3DTransformer(nn.Module):
    """query_embed -> (query_pos, query); initial reference points; decoder."""

    def __init__(self, decoder: Detr3DTransformerDecoder, num_query=900, code_size=10, num_reg_fcs=2):
        super().__init__()
        self.decoder = decoder
        self.embed_dims = decoder.embed_dims
        self.reference_points = nn.Linear(self.embed_dims, 3)
        self.query_embedding = nn.Embedding(num_query, self.embed_dims * 2)   # detr3d_head.py:97-99
        branches = []
        for _ in range(len(decoder.layers)):                                # detr3d_head.py:72-95
            fcs = []
            for _ in range(num_reg_fcs):
                fcs += [nn.Linear(self.embed_dims, self.embed_dims), nn.ReLU()]
            fcs.append(nn.Linear(self.embed_dims, code_size))
            branches.append(nn.Sequential(*fcs))
        self.reg_branches = nn.ModuleList(branches)
        nn.init.xavier_uniform_(self.reference_points.weight)
        nn.init.constant_(self.reference_points.bias, 0.0)

    def forward(self, mlvl_feats, img_metas, batch_size: int):
        query_embed = self.query_embedding.weight
        query_pos, query = torch.split(query_embed, self.embed_dims, dim=1)
        # the two halves are row-strided views of the (Q, 2C) embedding: make them dense ONCE per
        # forward instead of once per consumer (every fused kernel needs dense rows)
        query_pos, query = query_pos.contiguous(), query.contiguous()
        query_pos = query_pos.unsqueeze(0).expand(batch_size, -1, -1)
        query = query.unsqueeze(0).expand(batch_size, -1, -1)
        reference_points = fast_linear(query_pos, self.reference_points).sigmoid()   # NOT detached in layer 0
        inter_states, inter_refs = self.decoder(
            query.permute(1, 0, 2), mlvl_feats, query_pos.permute(1, 0, 2), reference_points,
            reg_branches=self.reg_branches, img_metas=img_metas)
        return inter_states, reference_points, inter_refs
