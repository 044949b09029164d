"""TEST INFRASTRUCTURE ONLY -- executes the *unmodified* reference code as the pin.

Loads Graph-DETR4D's cross-view attention classes straight from
``/root/reference`` (read-only) without mmcv/mmdet/mmdet3d being installed, by
putting ~40 lines of ``sys.modules`` shims in place (SURVEY.md section 8c,
Appendix C).  Nothing is copied: the reference sources are read where they lie
and executed.  This module only works in the build container; the GPU box has
no ``/root/reference``, which is why the outputs are frozen into
``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.

What is loaded (reference file:line):
  * ``feature_sampling``                 detr3d_transformer.py:397-438
  * ``Detr3DCrossAtten``                 detr3d_transformer.py:229-390
  * ``Detr3DCrossAttenV2``               detr3d_transformer.py:441-709
  * ``Detr3DTransformerDecoder``         detr3d_transformer.py:151-225   (its mmcv parent is the shim
        ``TransformerLayerSequence``: an nn.Module with an empty ``layers`` list the test fills)
  * ``Detr3DTransformer``                detr3d_transformer.py:45-147
  * ``Deform3DCrossAttn``                deform3d_cross_attn.py:33-339
        The shipped non-CUDA branch (deform3d_cross_attn.py:305-309) raises
        NameError (``sampling_locations`` is undefined).  ``variant="cpu"``
        applies ONE documented token substitution in that call
        (``sampling_locations`` -> ``reference_points_cam``) so the class runs
        on CPU through mmcv's public grid_sample formulation of multi-scale
        deformable attention (re-stated in ``_msda_pytorch`` below from mmcv
        1.x ``mmcv/ops/multi_scale_deform_attn.py::multi_scale_deformable_attn_pytorch``;
        mmcv is a third-party dependency that is NOT vendored in the reference
        and whose version the reference does not pin).

Only ``tests/`` and ``tests/golden/make_golden.py`` import this file.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("GD4D_REFERENCE_ROOT", "/root/reference")
REF_UTILS = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin", "models", "utils")
_PKG = "gd4d_refutils"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_UTILS, "detr3d_transformer.py"))


# --------------------------------------------------------------------------
# mmcv's public pure-PyTorch multi-scale deformable attention (grid_sample form)
# --------------------------------------------------------------------------
def _msda_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """mmcv 1.x ``multi_scale_deformable_attn_pytorch`` (third-party, restated).

    value (bs, num_keys, heads, dims); shapes (L,2) as (h,w);
    sampling_locations (bs, Q, heads, L, P, 2) in [0,1]; weights (bs,Q,heads,L,P).
    """
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, num_heads, num_levels, num_points, _ = sampling_locations.shape
    value_list = value.split([int(H_) * int(W_) for H_, W_ in value_spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for level, (H_, W_) in enumerate(value_spatial_shapes):
        H_, W_ = int(H_), int(W_)
        value_l_ = value_list[level].flatten(2).transpose(1, 2).reshape(
            bs * num_heads, embed_dims, H_, W_)
        sampling_grid_l_ = sampling_grids[:, :, :, level].transpose(1, 2).flatten(0, 1)
        sampling_value_l_ = F.grid_sample(
            value_l_, sampling_grid_l_, mode="bilinear", padding_mode="zeros",
            align_corners=False)
        sampling_value_list.append(sampling_value_l_)
    attention_weights = attention_weights.transpose(1, 2).reshape(
        bs * num_heads, 1, num_queries, num_levels * num_points)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights
              ).sum(-1).view(bs, num_heads * embed_dims, num_queries)
    return output.transpose(1, 2).contiguous()


class _MSDAFunction:
    """Stand-in for mmcv's CUDA autograd Function: same maths via grid_sample."""

    @staticmethod
    def apply(value, spatial_shapes, level_start_index, sampling_locations,
              attention_weights, im2col_step):
        return _msda_pytorch(value, spatial_shapes, sampling_locations, attention_weights)


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


def _xavier_init(module, gain=1, bias=0, distribution="normal"):
    if hasattr(module, "weight") and module.weight is not None:
        if distribution == "uniform":
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__gd4d_shim__ = True
        sys.modules[name] = m
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_mod(parent), leaf, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_REGS = {}


def install_shims():
    """Install import shims for mmcv / mmdet / mmdet3d (idempotent)."""
    if _REGS:
        return _REGS
    attention = _Registry("attention")
    tls = _Registry("transformer_layer_sequence")
    transformer = _Registry("transformer")
    _REGS.update(ATTENTION=attention, TRANSFORMER_LAYER_SEQUENCE=tls, TRANSFORMER=transformer)

    class TransformerLayerSequence(_BaseModule):
        def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
            super().__init__(init_cfg)
            self.num_layers = num_layers
            self.layers = nn.ModuleList()

    class MultiScaleDeformableAttention(_BaseModule):
        pass

    _mod("mmcv")
    _mod("mmcv.cnn", xavier_init=_xavier_init, constant_init=_constant_init)
    _mod("mmcv.cnn.bricks")
    _mod("mmcv.cnn.bricks.registry", ATTENTION=attention, TRANSFORMER_LAYER_SEQUENCE=tls)
    _mod("mmcv.cnn.bricks.transformer",
         MultiScaleDeformableAttention=MultiScaleDeformableAttention,
         TransformerLayerSequence=TransformerLayerSequence,
         # mmcv builds the sequence from a config dict; the pin hands over an already-built module
         build_transformer_layer_sequence=lambda cfg, *a, **k: cfg if isinstance(cfg, nn.Module) else None)
    _mod("mmcv.runner")
    _mod("mmcv.runner.base_module", BaseModule=_BaseModule)
    _mod("mmcv.ops")
    _mod("mmcv.ops.multi_scale_deform_attn",
         MultiScaleDeformableAttnFunction=_MSDAFunction,
         multi_scale_deformable_attn_pytorch=_msda_pytorch)
    _mod("mmdet3d")
    _mod("mmdet3d.core")
    _mod("mmdet3d.core.bbox")
    _mod("mmdet3d.core.bbox.structures")
    _mod("mmdet3d.core.bbox.structures.utils", rotation_3d_in_axis=None)
    _mod("mmdet")
    _mod("mmdet.models")
    _mod("mmdet.models.utils")
    _mod("mmdet.models.utils.builder", TRANSFORMER=transformer)
    _mod("mmdet.models.utils.transformer")
    pkg = _mod(_PKG)
    pkg.__path__ = [REF_UTILS]
    return _REGS


def _load_file(modname, filename, patch=None):
    full = f"{_PKG}.{modname}"
    if full in sys.modules and patch is None:
        return sys.modules[full]
    path = os.path.join(REF_UTILS, filename)
    if patch is None:
        spec = importlib.util.spec_from_file_location(full, path)
        module = importlib.util.module_from_spec(spec)
        sys.modules[full] = module
        spec.loader.exec_module(module)
        return module
    src = open(path).read()
    src = patch(src)
    module = types.ModuleType(full + "_patched")
    module.__file__ = path
    module.__package__ = _PKG
    exec(compile(src, path, "exec"), module.__dict__)
    return module


_DC_CALL = "value_flatten, spatial_shapes, sampling_locations, attention_weights)"


def _patch_dc(src):
    # the one documented token substitution (deform3d_cross_attn.py:308-309)
    assert src.count(_DC_CALL) == 1, "reference source changed; re-audit the patch"
    return src.replace(_DC_CALL, _DC_CALL.replace("sampling_locations", "reference_points_cam"))


def load():
    """Return a namespace with the reference symbols.

    ``Deform3DCrossAttn``      -- unmodified class (runs its CUDA branch on a GPU
                                   through the shimmed MSDA function)
    ``Deform3DCrossAttnCPU``   -- same file with the one-token fix, runs on CPU
    """
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    install_shims()
    dc = _load_file("deform3d_cross_attn", "deform3d_cross_attn.py")
    dt = _load_file("detr3d_transformer", "detr3d_transformer.py")
    dc_cpu = _load_file("deform3d_cross_attn", "deform3d_cross_attn.py", patch=_patch_dc)
    ns = types.SimpleNamespace(
        feature_sampling=dt.feature_sampling,
        inverse_sigmoid=dt.inverse_sigmoid,
        Detr3DCrossAtten=dt.Detr3DCrossAtten,
        Detr3DCrossAttenV2=dt.Detr3DCrossAttenV2,
        Detr3DTransformerDecoder=dt.Detr3DTransformerDecoder,      # detr3d_transformer.py:151-225 (row a9)
        Detr3DTransformer=dt.Detr3DTransformer,                    # detr3d_transformer.py:45-147
        Deform3DCrossAttn=dc.Deform3DCrossAttn,
        Deform3DCrossAttnCPU=dc_cpu.Deform3DCrossAttn,
        msda_pytorch=_msda_pytorch,
        registries=_REGS,
    )
    return ns


# ------------------------------------------------------------------------------------------
# Detr3DHeadPE.position_embeding: executed unmodified as the pin
# ------------------------------------------------------------------------------------------
REF_HEAD_PE = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin", "models", "dense_heads", "detr3d_head_pe.py")


def load_position_embeding():
    """The reference's ``Detr3DHeadPE.position_embeding`` (dense_heads/detr3d_head_pe.py:427-491) as
    a plain function ``f(self, img_feats, img_metas, masks)``.  The method's FunctionDef is
    AST-extracted from the file where it lies and compiled unmodified; the file itself cannot be
    imported (mmcv / mmdet / mmdet3d are not installed).  Its one free name that is not numpy /
    torch, ``inverse_sigmoid`` (imported there from mmdet 2.x, un-vendored, un-pinned), is bound to
    the reference's OWN identical copy, detr3d_transformer.py:28-43."""
    import ast
    import numpy as np
    src = open(REF_HEAD_PE).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "Detr3DHeadPE":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "position_embeding":
                    fn = item
    assert fn is not None, "reference source changed: Detr3DHeadPE.position_embeding not found"
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "torch": torch, "inverse_sigmoid": load().inverse_sigmoid}
    exec(compile(mod, REF_HEAD_PE, "exec"), ns)
    return ns["position_embeding"]


REF_POS_ENC = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin", "models", "utils", "positional_encoding.py")


def load_fpe_block():
    """The feature position-embedding block of ``Detr3DHeadPE.forward``
    (dense_heads/detr3d_head_pe.py:510-553), executed from the reference's own source:

      * ``forward`` is AST-extracted and CUT right before ``query_embeds = ...`` (:556, the transformer
        call and the branches are not part of row f4); a ``return mlvl_feats, masks`` is appended.
        Nothing inside the kept statements is touched.
      * ``SELayer`` (:231-243) and ``SinePositionalEncoding3D`` (models/utils/positional_encoding.py:15-112)
        are AST-extracted class definitions, compiled unmodified with ``BaseModule`` bound to
        ``torch.nn.Module``-with-init_cfg (the shim of this file) and the registry decorator dropped.
      * ``position_embeding`` is the method loaded by ``load_position_embeding``.

    Returns a namespace with ``forward_fpe(self, mlvl_feats, img_metas)``, ``SELayer``,
    ``SinePositionalEncoding3D`` and ``position_embeding``; ``make_head`` builds the stub ``self`` with the
    sub-modules the reference constructs at :386-396."""
    import ast
    import math
    import types
    import numpy as np
    import torch.nn as nn
    import torch.nn.functional as F
    install_shims()
    ns = {"np": np, "torch": torch, "nn": nn, "F": F, "math": math, "BaseModule": _BaseModule}
    tree = ast.parse(open(REF_HEAD_PE).read())
    fwd = se = None
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "SELayer":
            se = node
        if isinstance(node, ast.ClassDef) and node.name == "Detr3DHeadPE":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "forward":
                    fwd = item
    assert fwd is not None and se is not None, "reference source changed"
    keep = []
    for st in fwd.body:
        seg = ast.get_source_segment(open(REF_HEAD_PE).read(), st) or ""
        if seg.startswith("query_embeds"):
            break
        keep.append(st)
    assert len(keep) < len(fwd.body), "cut point (query_embeds = ...) not found"
    ret = ast.parse("return mlvl_feats, masks").body[0]
    fwd.body = keep + [ret]
    fwd.name = "forward_fpe"
    fwd.decorator_list = []
    exec(compile(ast.fix_missing_locations(ast.Module(body=[se, fwd], type_ignores=[])), REF_HEAD_PE, "exec"), ns)
    tree2 = ast.parse(open(REF_POS_ENC).read())
    pe = next(n for n in tree2.body if isinstance(n, ast.ClassDef) and n.name == "SinePositionalEncoding3D")
    pe.decorator_list = []
    exec(compile(ast.Module(body=[pe], type_ignores=[]), REF_POS_ENC, "exec"), ns)
    ns["position_embeding"] = load_position_embeding()

    def make_head(embed_dims=256, depth_num=64, depth_start=1, pc_range=None, with_detach=True, num_feats=128,
                  seed=0):
        torch.manual_seed(seed)
        head = nn.Module()
        head.embed_dims, head.depth_num, head.depth_start = embed_dims, depth_num, depth_start
        head.pc_range, head.with_detach = pc_range, with_detach
        head.position_dim = 3 * depth_num
        head.position_encoder = nn.Sequential(                                     # :386-390
            nn.Conv2d(head.position_dim, embed_dims * 4, kernel_size=1, stride=1, padding=0), nn.ReLU(),
            nn.Conv2d(embed_dims * 4, embed_dims, kernel_size=1, stride=1, padding=0))
        head.adapt_pos3d = nn.Sequential(                                          # :391-395
            nn.Conv2d(embed_dims * 3 // 2, embed_dims * 4, kernel_size=1, stride=1, padding=0), nn.ReLU(),
            nn.Conv2d(embed_dims * 4, embed_dims, kernel_size=1, stride=1, padding=0))
        head.fpe = ns["SELayer"](embed_dims)                                       # :396
        head.positional_encoding = ns["SinePositionalEncoding3D"](num_feats=num_feats, normalize=True, offset=-0.5)
        head.position_embeding = types.MethodType(ns["position_embeding"], head)
        return head

    ns["make_head"] = make_head
    return types.SimpleNamespace(**{k: ns[k] for k in ("forward_fpe", "SELayer", "SinePositionalEncoding3D",
                                                      "position_embeding", "make_head")})


# ------------------------------------------------------------------------------------------
# HungarianAssigner3D: the reference's own class, executed unmodified
# ------------------------------------------------------------------------------------------
REF_BBOX = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin", "core", "bbox")


def load_hungarian_assigner(focal_loss_cost_cls):
    """``HungarianAssigner3D`` (core/bbox/assigners/hungarian_assigner_3d.py:25-145), ``BBox3DL1Cost``
    (core/bbox/match_costs/match_cost.py:6-28) and ``normalize_bbox`` (core/bbox/util.py:38-57),
    AST-extracted from the files where they lie and compiled unmodified.  mmdet (un-vendored,
    un-pinned) supplies four names the class needs: ``BaseAssigner`` (abstract base -> ``object``),
    ``AssignResult`` (a record -> namedtuple with the same field order), ``build_match_cost`` (a
    registry lookup -> a 3-entry dict) and ``FocalLossCost`` -- the one with arithmetic in it, passed
    in by the caller as a restatement of mmdet 2.x's published formula.
    Returns (HungarianAssigner3D, BBox3DL1Cost, normalize_bbox)."""
    import ast
    import collections
    import numpy as np

    def extract(path, names):
        tree = ast.parse(open(path).read())
        body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
        assert len(body) == len(names), f"reference source changed: {names} not all found in {path}"
        for n in body:
            n.decorator_list = []                              # registry / array-converter decorators
        return ast.Module(body=body, type_ignores=[])

    AssignResult = collections.namedtuple("AssignResult", ["num_gts", "gt_inds", "max_overlaps", "labels"])

    class _IoUCost:                                            # weight 0.0 in every config, never called
        def __init__(self, weight=0.0, **kw):
            self.weight = weight

    ns_util = {"torch": torch}
    exec(compile(extract(os.path.join(REF_BBOX, "util.py"), ["normalize_bbox"]), "util.py", "exec"), ns_util)
    ns_cost = {"torch": torch}
    exec(compile(extract(os.path.join(REF_BBOX, "match_costs", "match_cost.py"), ["BBox3DL1Cost"]),
                 "match_cost.py", "exec"), ns_cost)
    table = {"FocalLossCost": focal_loss_cost_cls, "BBox3DL1Cost": ns_cost["BBox3DL1Cost"], "IoUCost": _IoUCost}

    def build_match_cost(cfg):
        cfg = dict(cfg)
        return table[cfg.pop("type")](**cfg)

    from scipy.optimize import linear_sum_assignment
    ns = {"torch": torch, "BaseAssigner": object, "AssignResult": AssignResult, "build_match_cost": build_match_cost,
          "normalize_bbox": ns_util["normalize_bbox"], "linear_sum_assignment": linear_sum_assignment, "np": np}
    exec(compile(extract(os.path.join(REF_BBOX, "assigners", "hungarian_assigner_3d.py"), ["HungarianAssigner3D"]),
                 "hungarian_assigner_3d.py", "exec"), ns)
    return ns["HungarianAssigner3D"], ns_cost["BBox3DL1Cost"], ns_util["normalize_bbox"]
