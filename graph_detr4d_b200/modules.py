"""Drop-in attention modules: same class names, constructor keys, forward
signature, parameter (state-dict) names and ``init_weight()`` as the reference's
mmcv ``ATTENTION``-registered classes, with the sampling core replaced by the
fused sm_100a kernels.

  Detr3DCrossAtten   <- projects/mmdet3d_plugin/models/utils/detr3d_transformer.py:229-390
  Deform3DCrossAttn  <- projects/mmdet3d_plugin/models/utils/deform3d_cross_attn.py:33-339
  Detr3DCrossAttenV2 <- projects/mmdet3d_plugin/models/utils/detr3d_transformer.py:441-709

What stays in torch (cuBLAS library GEMMs, tiny): the weight/offset generator
Linears, value_proj, output_proj, position_encoder.  What moved into ONE kernel
launch per layer: point de-normalisation, 3D offsets, lidar2img projection,
depth / in-image mask, softmax / sigmoid, 4-level bilinear sampling, the weighted
reduction over levels, points and cameras.

Differences from the reference, all deliberate:
  * feature maps are packed to channel-last ONCE per forward and shared by the
    decoder layers (the reference re-flattens them in every layer,
    deform3d_cross_attn.py:264-269); ``value`` may also be a ``PackedFeatures``.
  * ``lidar2img`` is converted/uploaded once per distinct set of matrices, not
    once per layer (detr3d_transformer.py:398-402).
  * ``residual`` is honoured (the reference leaves ``inp_residual`` undefined when
    it is not None: latent NameError at detr3d_transformer.py:363-364).
  * Deform3DCrossAttn with batch size > 1: the reference's ``query.repeat(N,1,1)``
    (deform3d_cross_attn.py:277) pairs attention logits of sample ``i % B`` with
    image ``i = b*N+n``; that is only self-consistent for B == 1 (every config uses
    samples_per_gpu=1).  The module therefore RAISES for B > 1 (as Detr3DCrossAttenV2 does),
    unless built with ``allow_batched=True`` (new key), in which case every sample uses its
    own logits -- NOT what the reference computes for B > 1 (the oracle reproduces the
    reference's pairing and is pinned to it: tests/test_oracle_vs_reference.py).
  * ``feature_dtype='bf16'`` (new): keep the packed / value-projected maps in bf16
    (fp32 accumulation in the kernel).
  * Deform3DCrossAttn ``value_proj_mode`` (new): ``'fused'`` (default) never runs
    value_proj over the pixels.  By linearity sum_s w_s (W f_s + b) = W sum_s w_s f_s
    + b sum_s w_s, so each head gathers all C raw channels (kernel "wide" mode) and
    W_v's head slice is applied to the (B,Hh,Q,C) result with one tiny batched GEMM.
    ``'dense'`` reproduces the reference's op boundary (value_proj GEMM over every
    pixel, then mmcv-layout head-slice sampling); it is used automatically when the
    channel count does not fit the wide kernel.
"""
from __future__ import annotations

import math
import warnings
import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused, ops
from .glue import can_defer_bias, cat_linear, fast_layer_norm, fast_linear
from .ops import MODE_A, MODE_C, MODE_V2, PackedFeatures, XViewConfig

import os as _os
import sys as _sys

_PACKED_GEN = _os.environ.get("GD4D_PACKED_GEN", "1") != "0"    # A/B switch for measurements


def _real_mmcv_available() -> bool:
    m = _sys.modules.get("mmcv")
    if m is not None and getattr(m, "__gd4d_shim__", False):
        return False          # the test-only import shim of oracle/ref_loader.py, not mmcv
    try:
        import mmcv.cnn.bricks.registry  # noqa: F401
        import mmcv.runner.base_module  # noqa: F401
        return True
    except Exception:
        return False


HAVE_MMCV = _real_mmcv_available()
if HAVE_MMCV:  # real mmcv present -> register into ITS registry so configs build our classes
    from mmcv.cnn.bricks.registry import ATTENTION  # type: ignore
    from mmcv.runner.base_module import BaseModule  # type: ignore
else:  # mmcv is not installed in this image: same decorator API, local registry
    class _LocalRegistry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                key = name or cls.__name__
                if key in self.module_dict and not force:
                    raise KeyError(f"{key} is already registered in {self.name}")
                self.module_dict[key] = cls
                return cls
            return deco(module) if module is not None else deco

        def get(self, key):
            return self.module_dict.get(key)

        def build(self, cfg):
            cfg = dict(cfg)
            return self.module_dict[cfg.pop("type")](**cfg)

    ATTENTION = _LocalRegistry("attention")

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg


def build_attention(cfg):
    """mmcv.cnn.bricks.transformer.build_attention equivalent for the local registry."""
    if HAVE_MMCV:
        from mmcv.cnn.bricks.transformer import build_attention as _b  # type: ignore
        return _b(cfg)
    return ATTENTION.build(cfg)


def inverse_sigmoid(x, eps=1e-5, clamp_max=False):
    """detr3d_transformer.py:28-43 / deform3d_cross_attn.py:16-31."""
    x = x.clamp(min=0, max=1)
    if clamp_max:
        x1, x2 = x.clamp(min=eps, max=1), (1 - x).clamp(min=eps, max=1)
    else:
        x1, x2 = x.clamp(min=eps), (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


def _xavier_uniform_(lin: nn.Linear, bias: float = 0.0):
    nn.init.xavier_uniform_(lin.weight, gain=1)
    nn.init.constant_(lin.bias, bias)


def _constant_(lin: nn.Linear, val: float, bias: float = 0.0):
    nn.init.constant_(lin.weight, val)
    nn.init.constant_(lin.bias, bias)


def _feature_dtype(name) -> Optional[torch.dtype]:
    if name in (None, "keep"):
        return None
    if name in ("bf16", "bfloat16", torch.bfloat16):
        return torch.bfloat16
    if name in ("fp32", "float32", torch.float32):
        return torch.float32
    raise ValueError(f"feature_dtype must be None, 'fp32' or 'bf16', got {name!r}")


# ----------------------------------------------------------------------------------------
# per-forward caches shared by the decoder layers
# ----------------------------------------------------------------------------------------
class _PackCache:
    """Single-entry cache: the 6 decoder layers receive the SAME python list of
    feature tensors (detr3d_transformer.py:140-147), so pack it once."""

    def __init__(self):
        self._refs = None
        self._versions = None
        self._dtype = None
        self._grad = None
        self._packed = None

    @staticmethod
    def _grad_state(value):
        # a pack made under no_grad has no gradient sink: it must not serve a later training forward
        return (torch.is_grad_enabled(), tuple(bool(v.requires_grad) for v in value))

    def get(self, value: Sequence[torch.Tensor], dtype) -> PackedFeatures:
        if self._refs is not None and len(self._refs) == len(value) and self._dtype == dtype and \
                all(r() is v for r, v in zip(self._refs, value)) and \
                self._versions == [v._version for v in value] and self._grad == self._grad_state(value):
            return self._packed
        packed = ops.pack_features(value, dtype)
        self._refs = [weakref.ref(v) for v in value]
        self._versions = [v._version for v in value]
        self._dtype = dtype
        self._grad = self._grad_state(value)
        self._packed = packed
        return packed

    def clear(self):
        self.__init__()


class _Lidar2ImgCache:
    """One persistent (B,N,4,4) fp32 device tensor per shape/device, refreshed IN PLACE
    only when the matrices change: hoists the reference's per-layer list->numpy->tensor
    H2D copy (detr3d_transformer.py:398-402) to once per distinct sample, and gives
    CUDA-graph replays a static address to read."""

    static = False          # set by lidar2img_device(): keep ONE address (graph capture), refresh in place

    def __init__(self):
        self._arr = None
        self._key = None
        self._tensor = None

    def get(self, img_metas, device) -> torch.Tensor:
        arr = np.asarray([m["lidar2img"] for m in img_metas]).astype(np.float32)
        key = (arr.shape, torch.device(device))
        if self._tensor is not None and self._key == key:
            if not np.array_equal(arr, self._arr):
                if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("lidar2img changed during CUDA-graph capture; call "
                                       "lidar2img_device(img_metas, device) before capturing")
                self._arr = arr
                if self.static or not torch.is_grad_enabled():
                    with torch.no_grad():                      # CUDA-graph replays read this address
                        self._tensor.copy_(torch.from_numpy(arr))
                else:
                    # an earlier forward may have saved the old tensor for its backward (two forwards
                    # before one backward: teacher/student, multi-sample losses): never overwrite it
                    self._tensor = torch.from_numpy(arr).to(device)
            return self._tensor
        self._arr, self._key = arr, key
        self._tensor = torch.from_numpy(arr).to(device)
        return self._tensor


_PACK_CACHE = _PackCache()
_L2I_CACHE = _Lidar2ImgCache()


def clear_pack_cache():
    _PACK_CACHE.clear()


def clear_caches():
    from . import ops as _ops
    _ops.clear_scratch_pool()
    _PACK_CACHE.clear()
    _L2I_CACHE.__init__()
    _L2I_CACHE.static = False


def lidar2img_device(img_metas, device) -> torch.Tensor:
    """Upload / refresh the STATIC lidar2img buffer (call outside CUDA-graph capture): from now on the
    buffer keeps its address and is refreshed in place, which is what graph replays need."""
    _L2I_CACHE.static = True
    return _L2I_CACHE.get(img_metas, device)


def _get_packed(value, dtype) -> PackedFeatures:
    if isinstance(value, PackedFeatures):
        return value
    if isinstance(value, torch.Tensor):
        raise TypeError("value must be the list of (B,N,C,H,W) feature maps (or PackedFeatures)")
    return _PACK_CACHE.get(list(value), dtype)


def _img_hw(img_metas):
    shp = img_metas[0]["img_shape"][0]          # detr3d_transformer.py:419-420: sample 0, camera 0
    return float(shp[0]), float(shp[1])


def _run_position_encoder(seq: nn.Sequential, x):
    """Linear-LN-ReLU-Linear-LN-ReLU; on CUDA fp32 every LN+ReLU pair is one fused launch."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (isinstance(nxt, nn.LayerNorm) and can_defer_bias(x, m) and nxt.normalized_shape == (m.out_features,)
                    and fused.can_fuse_layernorm(x.new_empty(0, m.out_features), nxt)):
                # Linear -> LN (-> ReLU): GEMM without epilogue, bias + LN (+ ReLU) in one launch
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                x = fused.add_layernorm(fast_linear(x, m, add_bias=False), nxt, relu=relu, xbias=m.bias)
                i += 2 if relu else 1
            else:
                x = fast_linear(x, m)
        elif isinstance(m, nn.LayerNorm):
            relu_next = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            if relu_next and fused.can_fuse_layernorm(x, m):
                x = fused.add_layernorm(x, m, relu=True)
                i += 1
            else:
                x = fast_layer_norm(x, m)
        else:
            x = m(x)
        i += 1
    return x


def _output_proj(x, lin: nn.Linear, dropout: nn.Module, defer_bias: bool):
    """dropout(lin(x)) -> (tensor, pending_bias).  The bias is left to the caller (one launch less:
    fp32 cuBLAS runs the bias epilogue as its own kernel) only when asked and dropout is inactive."""
    active = isinstance(dropout, nn.Dropout) and dropout.training and dropout.p > 0
    if defer_bias and fused.ENABLED and not active and can_defer_bias(x, lin):
        return fast_linear(x, lin, add_bias=False), lin.bias
    return dropout(fast_linear(x, lin)), None


def _inv_sigmoid(x, clamp_max=False):
    """inverse_sigmoid as one launch on the path (CUDA fp32); the op-by-op torch version otherwise."""
    if fused.ENABLED and x.is_cuda and x.dtype == torch.float32:
        return fused.inverse_sigmoid(x, 1e-5, clamp_max)
    return inverse_sigmoid(x, clamp_max=clamp_max)


def _position_encoder(in_dims, embed_dims):
    return nn.Sequential(
        nn.Linear(in_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True),
        nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True))


# ----------------------------------------------------------------------------------------
# variant A
# ----------------------------------------------------------------------------------------
@ATTENTION.register_module(force=True)
class Detr3DCrossAtten(BaseModule):
    """DETR3D centre-point cross-view attention (detr3d_transformer.py:229-390)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=5, num_cams=6,
                 im2col_step=64, pc_range=None, dropout=0.1, norm_cfg=None, init_cfg=None,
                 batch_first=False, feature_dtype=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, "
                             f"but got {embed_dims} and {num_heads}")
        if not _is_power_of_2(embed_dims // num_heads):
            warnings.warn("You'd better set embed_dims in MultiScaleDeformAttention to make the "
                          "dimension of each attention head a power of 2 which is more efficient "
                          "in our CUDA implementation.")
        self.norm_cfg = norm_cfg
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.num_cams = num_cams
        self.attention_weights = nn.Linear(embed_dims, num_cams * num_levels * num_points)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.position_encoder = _position_encoder(3, embed_dims)
        self.batch_first = batch_first
        self.feature_dtype = _feature_dtype(feature_dtype)
        self.init_weight()

    def init_weight(self):
        _constant_(self.attention_weights, 0.0, 0.0)
        _xavier_uniform_(self.output_proj, 0.0)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        out, inp_residual, pos_feat = self.forward_parts(
            query, key, value, residual, query_pos=query_pos, key_padding_mask=key_padding_mask,
            reference_points=reference_points, spatial_shapes=spatial_shapes,
            level_start_index=level_start_index, **kwargs)
        return out + inp_residual + pos_feat

    def forward_parts(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                      reference_points=None, spatial_shapes=None, level_start_index=None,
                      defer_bias=False, query_with_pos=None, **kwargs):
        """The three addends of ``forward`` -- dropout(output_proj(sampled)), the residual and the
        position feature -- so that a caller that owns the following LayerNorm (decoder.py) can
        fold the sum into it (one launch).  With ``defer_bias`` a 4th item is returned: the
        output_proj bias still to be added to the first addend (or None if it already was).
        ``query_with_pos``: ``query + query_pos`` if the caller already has it."""
        if key is None:
            key = query
        if value is None:
            value = key
        inp_residual = query if residual is None else residual
        if query_with_pos is not None:                      # caller already has query + query_pos (fused LN output)
            query = query_with_pos
        elif query_pos is not None:
            query = query + query_pos
        query = query.permute(1, 0, 2)                      # (B,Q,C)
        img_metas = kwargs["img_metas"]
        packed = _get_packed(value, self.feature_dtype)
        if packed.N != self.num_cams or len(packed.levels) != self.num_levels:
            raise ValueError(f"expected {self.num_cams} cams x {self.num_levels} levels, got "
                             f"{packed.N} x {len(packed.levels)}")
        logits = fast_linear(query, self.attention_weights)  # (B,Q,N*P*L) viewed (B,1,Q,N,P,L)
        img_h, img_w = _img_hw(img_metas)
        cfg = XViewConfig(MODE_A, self.num_heads, self.num_points, tuple(self.pc_range), img_h, img_w)
        l2i = _L2I_CACHE.get(img_metas, query.device)
        out = ops.xview_attention(cfg, packed, reference_points, logits, lidar2img=l2i)   # (B,Q,C)
        out, pending = _output_proj(out.permute(1, 0, 2), self.output_proj, self.dropout, defer_bias)
        pos_feat = _run_position_encoder(self.position_encoder, _inv_sigmoid(reference_points)).permute(1, 0, 2)
        return (out, inp_residual, pos_feat, pending) if defer_bias else (out, inp_residual, pos_feat)


# ----------------------------------------------------------------------------------------
# variant V2 (registered by the reference, used by no config)
# ----------------------------------------------------------------------------------------
@ATTENTION.register_module(force=True)
class Detr3DCrossAttenV2(BaseModule):
    """Deformable-DETR style 2D offsets around the projected centre
    (detr3d_transformer.py:441-709).  Like the reference it needs
    num_points == num_levels and batch size 1."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=5, num_cams=6,
                 im2col_step=64, pc_range=None, dropout=0.1, norm_cfg=None, init_cfg=None,
                 batch_first=False, feature_dtype=None):
        super().__init__(init_cfg)
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, "
                             f"but got {embed_dims} and {num_heads}")
        if not _is_power_of_2(embed_dims // num_heads):
            warnings.warn("You'd better set embed_dims in MultiScaleDeformAttention to make the "
                          "dimension of each attention head a power of 2 which is more efficient "
                          "in our CUDA implementation.")
        self.norm_cfg = norm_cfg
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.num_cams = num_cams
        self.attention_weights = nn.Linear(embed_dims, num_cams * num_heads * num_levels * num_points)
        self.sampling_offsets = nn.Linear(embed_dims, num_cams * num_heads * num_levels * num_points * 2)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.position_encoder = _position_encoder(3, embed_dims)
        self.batch_first = batch_first
        self.feature_dtype = _feature_dtype(feature_dtype)
        self.init_weight()

    def init_weight(self):
        _constant_(self.sampling_offsets, 0.0)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(
            1, self.num_heads, 1, 1, 2).repeat(self.num_cams, 1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        _constant_(self.attention_weights, 0.0, 0.0)
        _xavier_uniform_(self.output_proj, 0.0)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        out, inp_residual, pos_feat = self.forward_parts(
            query, key, value, residual, query_pos=query_pos, key_padding_mask=key_padding_mask,
            reference_points=reference_points, spatial_shapes=spatial_shapes,
            level_start_index=level_start_index, **kwargs)
        return out + inp_residual + pos_feat

    def forward_parts(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                      reference_points=None, spatial_shapes=None, level_start_index=None,
                      defer_bias=False, query_with_pos=None, **kwargs):
        """The three addends of ``forward`` -- dropout(output_proj(sampled)), the residual and the
        position feature -- so that a caller that owns the following LayerNorm (decoder.py) can
        fold the sum into it (one launch).  With ``defer_bias`` a 4th item is returned: the
        output_proj bias still to be added to the first addend (or None if it already was).
        ``query_with_pos``: ``query + query_pos`` if the caller already has it."""
        if key is None:
            key = query
        if value is None:
            value = key
        inp_residual = query if residual is None else residual
        if query_with_pos is not None:                      # caller already has query + query_pos (fused LN output)
            query = query_with_pos
        elif query_pos is not None:
            query = query + query_pos
        query = query.permute(1, 0, 2)
        img_metas = kwargs["img_metas"]
        packed = _get_packed(value, self.feature_dtype)
        if packed.B != 1:
            raise ValueError("Detr3DCrossAttenV2 broadcasts (B*N) against N: batch size must be 1 "
                             "(detr3d_transformer.py:700)")
        logits = fast_linear(query, self.attention_weights)      # (B,Q,N*Hh*L*P)
        offsets = fast_linear(query, self.sampling_offsets)      # (B,Q,N*Hh*L*P*2)
        img_h, img_w = _img_hw(img_metas)
        cfg = XViewConfig(MODE_V2, self.num_heads, self.num_points, tuple(self.pc_range), img_h, img_w)
        l2i = _L2I_CACHE.get(img_metas, query.device)
        out = ops.xview_attention(cfg, packed, reference_points, logits, offsets, None, l2i)   # (B,Q,C)
        out, pending = _output_proj(out.permute(1, 0, 2), self.output_proj, self.dropout, defer_bias)
        pos_feat = _run_position_encoder(self.position_encoder, _inv_sigmoid(reference_points)).permute(1, 0, 2)
        return (out, inp_residual, pos_feat, pending) if defer_bias else (out, inp_residual, pos_feat)


# ----------------------------------------------------------------------------------------
# variant C (Graph-DETR4D)
# ----------------------------------------------------------------------------------------
@ATTENTION.register_module(force=True)
class Deform3DCrossAttn(BaseModule):
    """Graph-DETR4D 3D-offset cross-view attention (deform3d_cross_attn.py:33-339)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=5, num_cams=6,
                 im2col_step=64, pc_range=None, dropout=0.1, norm_cfg=None, init_cfg=None,
                 batch_first=False, fix_offset=False, depth_encode=False, feature_dtype=None,
                 value_proj_mode="fused", allow_batched=False):
        super().__init__(init_cfg)
        if value_proj_mode not in ("fused", "dense"):
            raise ValueError("value_proj_mode must be 'fused' or 'dense'")
        self.value_proj_mode = value_proj_mode
        self.allow_batched = bool(allow_batched)
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, "
                             f"but got {embed_dims} and {num_heads}")
        if not _is_power_of_2(embed_dims // num_heads):
            warnings.warn("You'd better set embed_dims in MultiScaleDeformAttention to make the "
                          "dimension of each attention head a power of 2 which is more efficient "
                          "in our CUDA implementation.")
        self.norm_cfg = norm_cfg
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.fix_offset = fix_offset
        self.depth_encode = depth_encode
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.num_cams = num_cams
        self.cam_attention_weights = nn.Linear(embed_dims, num_cams)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.position_encoder = _position_encoder(4 if depth_encode else 3, embed_dims)
        self.batch_first = batch_first
        self.deform_sampling_offsets = nn.Linear(embed_dims, num_heads * 1 * num_points * 3)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.feature_dtype = _feature_dtype(feature_dtype)
        self.init_weight()
        if self.fix_offset:
            self.deform_sampling_offsets.weight.requires_grad = False
            self.deform_sampling_offsets.bias.requires_grad = False

    def init_weight(self):
        _constant_(self.cam_attention_weights, 0.0, 0.0)
        _xavier_uniform_(self.output_proj, 0.0)
        _constant_(self.deform_sampling_offsets, 0.0)
        # ring of directions x (p+1): deform3d_cross_attn.py:138-148
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin(), thetas.cos()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(
            self.num_heads, 1, 1, 3).repeat(1, 1, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.deform_sampling_offsets.bias.data = grid_init.view(-1)
        _constant_(self.attention_weights, 0.0, 0.0)
        _xavier_uniform_(self.value_proj, 0.0)

    _gen_pack = None
    _gen_pack_fresh = False

    def _generator_linears(self):
        return (self.attention_weights, self.deform_sampling_offsets, self.cam_attention_weights)

    def generator_pack_slots(self):
        """(destination views, source parameters) of this module's packed generator weights
        [attention_weights | deform_sampling_offsets | cam_attention_weights | zero pad].  A caller
        that copies sources into destinations (decoder.py: ONE multi-tensor copy for all layers)
        and then sets ``_gen_pack_fresh`` lets the next forward skip its own two concatenations."""
        lins = self._generator_linears()
        n = sum(l.out_features for l in lins)
        width = (n + 3) // 4 * 4
        w0 = lins[0].weight
        if self._gen_pack is None or self._gen_pack[0].device != w0.device:
            with torch.no_grad():
                self._gen_pack = (w0.new_zeros(width, w0.shape[1]), w0.new_zeros(width))
        wc, bc = self._gen_pack
        dsts, srcs, o = [], [], 0
        for l in lins:
            dsts += [wc[o:o + l.out_features], bc[o:o + l.out_features]]
            srcs += [l.weight, l.bias]
            o += l.out_features
        return dsts, srcs

    def _use_wide(self, packed: PackedFeatures) -> bool:
        row_bytes = packed.C * packed.levels[0].element_size()
        return self.value_proj_mode == "fused" and row_bytes in (512, 1024)

    def project_values(self, packed: PackedFeatures) -> List[torch.Tensor]:
        """value_proj over every pixel, level by level, straight on the channel-last
        maps (deform3d_cross_attn.py:278-280); output stays channel-last."""
        w, b = self.value_proj.weight, self.value_proj.bias
        vals = []
        for v in packed.differentiable_levels():
            if v.dtype != w.dtype:
                vals.append(F.linear(v, w.to(v.dtype), b.to(v.dtype)))
            else:
                vals.append(F.linear(v, w, b))
        return vals

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        out, inp_residual, pos_feat = self.forward_parts(
            query, key, value, residual, query_pos=query_pos, key_padding_mask=key_padding_mask,
            reference_points=reference_points, spatial_shapes=spatial_shapes,
            level_start_index=level_start_index, **kwargs)
        return out + inp_residual + pos_feat

    def forward_parts(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                      reference_points=None, spatial_shapes=None, level_start_index=None,
                      defer_bias=False, query_with_pos=None, **kwargs):
        """The three addends of ``forward`` -- dropout(output_proj(sampled)), the residual and the
        position feature -- so that a caller that owns the following LayerNorm (decoder.py) can
        fold the sum into it (one launch).  With ``defer_bias`` a 4th item is returned: the
        output_proj bias still to be added to the first addend (or None if it already was).
        ``query_with_pos``: ``query + query_pos`` if the caller already has it."""
        if key is None:
            key = query
        if value is None:
            value = key
        inp_residual = query if residual is None else residual
        if query_with_pos is not None:                      # caller already has query + query_pos (fused LN output)
            query = query_with_pos
        elif query_pos is not None:
            query = query + query_pos
        query = query.permute(1, 0, 2)                      # (B,Q,C)
        img_metas = kwargs["img_metas"]
        packed = _get_packed(value, self.feature_dtype)
        if packed.N != self.num_cams or len(packed.levels) != self.num_levels:
            raise ValueError(f"expected {self.num_cams} cams x {self.num_levels} levels, got "
                             f"{packed.N} x {len(packed.levels)}")
        if packed.B != 1 and not self.allow_batched:
            raise ValueError(
                "Deform3DCrossAttn: batch size > 1.  The reference pairs image b*N+n with the attention "
                "logits of sample (b*N+n) % B (query.repeat(N,1,1), deform3d_cross_attn.py:277), which is "
                "only self-consistent for B == 1 (all configs use samples_per_gpu=1).  Build the module "
                "with allow_batched=True to give every sample its own logits instead.")
        img_h, img_w = _img_hw(img_metas)
        l2i = _L2I_CACHE.get(img_metas, query.device)
        # The three generator Linears read the same query: on the CUDA fp32 path they are ONE GEMM over
        # the concatenated weights, and the kernels read / write their column blocks in place (gen_stride).
        packed_gen = fused.ENABLED and _PACKED_GEN and can_defer_bias(query, self.attention_weights)
        if packed_gen:
            n_attn = self.num_heads * self.num_levels * self.num_points
            n_off = self.num_heads * self.num_points * 3
            layout = ops.GenLayout(cam=n_attn + n_off, offsets=n_attn, attn=0,
                                   width=(n_attn + n_off + self.num_cams + 3) // 4 * 4)
            fresh, self._gen_pack_fresh = self._gen_pack_fresh, False
            gen = cat_linear(query, self._generator_linears(), layout.width,
                             packed=self._gen_pack if fresh else None)                   # (B,Q,width)

            def sample(cfg, values=None):
                return ops.xview_attention_gen(cfg, packed, reference_points, gen, layout, l2i, values=values)
        else:
            cam_logits = fast_linear(query, self.cam_attention_weights)    # (B,Q,N); kernel reads it as view(B,N,Q)
            offsets = fast_linear(query, self.deform_sampling_offsets)     # (B,Q,Hh*P*3)
            logits = fast_linear(query, self.attention_weights)            # (B,Q,Hh*L*P)

            def sample(cfg, values=None):
                return ops.xview_attention(cfg, packed, reference_points, logits, offsets, cam_logits, l2i,
                                           values=values)
        if self._use_wide(packed):
            cfg = XViewConfig(MODE_C, self.num_heads, self.num_points, tuple(self.pc_range), img_h, img_w,
                              wide=True)
            agg, wsum = sample(cfg)
            Hh, Ch = self.num_heads, self.embed_dims // self.num_heads
            wv = self.value_proj.weight.view(Hh, Ch, self.embed_dims)       # out channel = h*Ch + c
            # agg is head-major (B,Hh,Q,C): one strided-batched GEMM, no transpose copy of the 7 MB aggregate
            # W_v agg + b wsum as GEMMs only: the bias term is a K=1 batched GEMM folded into the main
            # one by baddbmm, so neither forward nor backward needs elementwise / reduction launches
            Bq, Qn = agg.shape[0], agg.shape[2]
            wt = wv.transpose(1, 2).unsqueeze(0).expand(Bq, Hh, self.embed_dims, Ch).reshape(Bq * Hh, self.embed_dims, Ch)
            bv = self.value_proj.bias.view(1, Hh, 1, Ch).expand(Bq, Hh, 1, Ch).reshape(Bq * Hh, 1, Ch)
            out = torch.baddbmm(torch.bmm(wsum.reshape(Bq * Hh, Qn, 1), bv),
                                agg.reshape(Bq * Hh, Qn, self.embed_dims), wt)   # (B*Hh,Q,Ch)
            out = out.view(Bq, Hh, Qn, Ch).permute(0, 2, 1, 3).flatten(2)       # (B,Q,C)
        else:
            cfg = XViewConfig(MODE_C, self.num_heads, self.num_points, tuple(self.pc_range), img_h, img_w)
            out = sample(cfg, self.project_values(packed))                 # (B,Q,C)
        out, pending = _output_proj(out, self.output_proj, self.dropout, defer_bias)
        out = out.permute(1, 0, 2)
        r3d = reference_points
        if self.depth_encode:
            depth = (r3d[..., 0:1] ** 2 + r3d[..., 1:2] ** 2) ** 0.5
            r3d = torch.cat([r3d, depth], dim=-1)
        pos_feat = _run_position_encoder(self.position_encoder, _inv_sigmoid(r3d, clamp_max=True)).permute(1, 0, 2)
        return (out, inp_residual, pos_feat, pending) if defer_bias else (out, inp_residual, pos_feat)
