"""Host logic of the drop-in modules that needs no GPU."""
import inspect

import numpy as np
import pytest
import torch

import graph_detr4d_b200 as g
from graph_detr4d_b200 import modules, synthetic as syn
from tests import helpers as H
from tests.test_oracle_golden import load_golden


def test_registry_builds_by_type_name():
    for name in ("Detr3DCrossAtten", "Deform3DCrossAttn", "Detr3DCrossAttenV2"):
        assert g.ATTENTION.get(name) is not None
    m = g.build_attention(dict(type="Deform3DCrossAttn", embed_dims=256, num_heads=8, num_levels=4,
                               num_points=4, num_cams=12, pc_range=syn.PC_RANGE, dropout=0.1,
                               batch_first=False))               # mmcv injects batch_first
    assert isinstance(m, g.Deform3DCrossAttn) and hasattr(m, "init_weight")


def test_constructor_errors_and_warnings():
    with pytest.raises(ValueError):
        g.Detr3DCrossAtten(embed_dims=100, num_heads=8, pc_range=syn.PC_RANGE)
    with pytest.warns(UserWarning):
        g.Detr3DCrossAtten(embed_dims=96, num_heads=8, pc_range=syn.PC_RANGE)   # head dim 12


def test_forward_signature_matches_reference_contract():
    want = ["query", "key", "value", "residual", "query_pos", "key_padding_mask", "reference_points",
            "spatial_shapes", "level_start_index", "kwargs"]
    for cls in (g.Detr3DCrossAtten, g.Deform3DCrossAttn, g.Detr3DCrossAttenV2):
        assert list(inspect.signature(cls.forward).parameters)[1:] == want


@pytest.mark.parametrize("variant,cls", [("A", "Detr3DCrossAtten"), ("C", "Deform3DCrossAttn"),
                                         ("V2", "Detr3DCrossAttenV2")])
def test_golden_state_dict_loads_strictly(variant, cls):
    gd = load_golden(variant)
    N = 6 * gd["T"]
    kw = dict(embed_dims=64, num_heads=2, num_levels=4, num_points=1 if variant == "A" else 4,
              num_cams=N, pc_range=syn.PC_RANGE)
    m = getattr(g, cls)(**kw)
    missing, unexpected = m.load_state_dict(gd["sd"], strict=True)
    assert not missing and not unexpected


def test_offset_bias_ring_init():
    m = g.Deform3DCrossAttn(num_cams=12, num_points=4, pc_range=syn.PC_RANGE)
    b = m.deform_sampling_offsets.bias.view(8, 4, 3)
    assert torch.allclose(b[0, 0], torch.tensor([1.0, 0.0, 1.0]))
    assert torch.allclose(b[:, 3], b[:, 0] * 4)
    assert m.deform_sampling_offsets.weight.abs().max() == 0
    m2 = g.Deform3DCrossAttn(num_cams=12, num_points=4, pc_range=syn.PC_RANGE, fix_offset=True)
    assert not m2.deform_sampling_offsets.weight.requires_grad


def test_cpu_tensors_are_refused_not_emulated():
    sc = H.scene(B=1, T=1, Q=4)
    m = g.Detr3DCrossAtten(num_cams=6, num_points=1, pc_range=syn.PC_RANGE).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(sc["query"], None, sc["feats"], query_pos=sc["query_pos"], reference_points=sc["ref"],
          img_metas=sc["metas"])


def test_product_never_imports_oracle():
    import os, re
    root = os.path.dirname(os.path.abspath(g.__file__))
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_lidar2img_cache_reuses_upload_only_for_equal_matrices():
    c = modules._Lidar2ImgCache()
    metas = syn.make_img_metas(1, 1)
    a = c.get(metas, torch.device("cpu"))
    b = c.get(syn.make_img_metas(1, 1), torch.device("cpu"))
    assert a is b and a.dtype == torch.float32 and tuple(a.shape) == (1, 6, 4, 4)
    metas[0]["lidar2img"][0] = metas[0]["lidar2img"][0] * 2
    old = a.clone()
    a2 = c.get(metas, torch.device("cpu"))
    # autograd may be recording: an earlier forward can have SAVED the old tensor for its backward
    # (two forwards before one backward), so changed matrices get a NEW tensor, the old one is intact
    assert a2 is not a and torch.equal(a, old)
    assert torch.equal(a2[0, 0], torch.as_tensor(metas[0]["lidar2img"][0].astype(np.float32)))
    # static mode (CUDA-graph replays read a fixed address) and no_grad: refreshed in place
    c.static = True
    metas[0]["lidar2img"][1] = metas[0]["lidar2img"][1] * 3
    a3 = c.get(metas, torch.device("cpu"))
    assert a3 is a2 and torch.equal(a3[0, 1], torch.as_tensor(metas[0]["lidar2img"][1].astype(np.float32)))
    c.static = False
    with torch.no_grad():
        metas[0]["lidar2img"][2] = metas[0]["lidar2img"][2] * 5
        assert c.get(metas, torch.device("cpu")) is a2


def test_pack_cache_key_includes_grad_state():
    """A pack made under no_grad has no gradient sink; it must not be served to a later training
    forward on the same (static) feature tensors."""
    assert modules._PackCache._grad_state([torch.zeros(1, requires_grad=True)]) == (True, (True,))
    with torch.no_grad():
        assert modules._PackCache._grad_state([torch.zeros(1, requires_grad=True)]) == (False, (True,))


def test_deform3d_refuses_batches_unless_allowed():
    import graph_detr4d_b200 as g
    m = g.Deform3DCrossAttn(num_cams=6, num_points=4, pc_range=syn.PC_RANGE)
    assert m.allow_batched is False
    assert g.Deform3DCrossAttn(num_cams=6, num_points=4, pc_range=syn.PC_RANGE, allow_batched=True).allow_batched


def test_synthetic_rig_valid_fraction():
    # Appendix B: ~0.18 of (query, camera) pairs project inside the image
    from oracle import xview_oracle as xo
    sc = H.scene(B=1, T=2, Q=2000)
    pts = xo.denormalize(sc["ref"], syn.PC_RANGE)
    uv, mask = xo.project(pts, sc["l2i"], 900, 1600)
    ok = mask[..., 0] & (uv[..., 0] > 0) & (uv[..., 0] < 1) & (uv[..., 1] > 0) & (uv[..., 1] < 1)
    assert 0.16 < ok.float().mean() < 0.20
