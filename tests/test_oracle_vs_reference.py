"""Pins the oracle restatement against the reference executed UNMODIFIED from
/root/reference (build container only; skipped where the tree is absent)."""
import contextlib
import io
import warnings

import pytest
import torch

from graph_detr4d_b200 import synthetic as syn
from oracle import ref_loader, xview_oracle as xo
from tests import helpers as H

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")]
warnings.filterwarnings("ignore", message="Default grid_sample")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 2)])
def test_feature_sampling_mask_bit_exact_and_samples_equal(ref, B, T):
    sc = H.scene(B=B, T=T, Q=300)
    r3d, sampled_ref, mask_ref = ref.feature_sampling(sc["feats"], sc["ref"], syn.PC_RANGE, sc["metas"])
    l2i = xo.lidar2img_tensor(sc["metas"], sc["ref"])
    sampled, mask = xo.feature_sampling_a(sc["feats"], sc["ref"], syn.PC_RANGE, l2i, 900, 1600)
    assert mask_ref.dtype == torch.bool and torch.equal(mask, mask_ref)
    assert 0.1 < mask.float().mean() < 0.3
    assert torch.equal(sampled, sampled_ref)


def test_explicit_matvec_equals_reference_matmul(ref):
    sc = H.scene(B=1, T=2, Q=2000)
    l2i = xo.lidar2img_tensor(sc["metas"], sc["ref"])
    pts = xo.denormalize(sc["ref"], syn.PC_RANGE)
    B, M = pts.shape[:2]
    N = l2i.size(1)
    hom = torch.cat((pts, torch.ones_like(pts[..., :1])), -1)
    cam = torch.matmul(l2i.view(B, N, 1, 4, 4).repeat(1, 1, M, 1, 1),
                       hom.view(B, 1, M, 4).repeat(1, N, 1, 1).unsqueeze(-1)).squeeze(-1)
    uv_ref = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * 1e-5)
    uv_ref[..., 0] /= 1600
    uv_ref[..., 1] /= 900
    uv, _ = xo.project(pts, l2i, 900, 1600)
    assert torch.equal(uv, uv_ref)


@pytest.mark.parametrize("B,T,P", [(1, 1, 1), (2, 1, 1), (1, 2, 3)])
def test_variant_a_module(ref, B, T, P):
    sc = H.scene(B=B, T=T, Q=150)
    torch.manual_seed(3)
    mod = ref.Detr3DCrossAtten(num_cams=sc["N"], num_points=P, pc_range=syn.PC_RANGE).eval()
    syn.randomize_generators(mod)
    feats = [f.clone().requires_grad_(True) for f in sc["feats"]]
    rp = sc["ref"].clone().requires_grad_(True)
    y_ref = mod(sc["query"], None, feats, query_pos=sc["query_pos"], reference_points=rp, img_metas=sc["metas"])
    y_ref.square().sum().backward()
    feats2 = [f.clone().requires_grad_(True) for f in sc["feats"]]
    rp2 = sc["ref"].clone().requires_grad_(True)
    y = xo.detr3d_cross_atten_forward(mod.state_dict(), sc["query"], feats2, sc["query_pos"], rp2,
                                      sc["metas"], syn.PC_RANGE)
    y.square().sum().backward()
    assert H.rel_err(y, y_ref) <= 2e-6
    assert H.rel_err(rp2.grad, rp.grad) <= 1e-5
    for a, b in zip(feats2, feats):
        assert H.rel_err(a.grad, b.grad) <= 1e-5


@pytest.mark.parametrize("T,P", [(1, 4), (2, 4), (2, 2)])
def test_variant_c_module(ref, T, P):
    sc = H.scene(B=1, T=T, Q=120)
    torch.manual_seed(4)
    mod = ref.Deform3DCrossAttnCPU(num_cams=sc["N"], num_points=P, pc_range=syn.PC_RANGE).eval()
    syn.randomize_generators(mod)
    feats = [f.clone().requires_grad_(True) for f in sc["feats"]]
    rp = sc["ref"].clone().requires_grad_(True)
    y_ref = _quiet(mod, sc["query"], None, feats, query_pos=sc["query_pos"], reference_points=rp,
                   img_metas=sc["metas"])
    y_ref.square().sum().backward()
    feats2 = [f.clone().requires_grad_(True) for f in sc["feats"]]
    rp2 = sc["ref"].clone().requires_grad_(True)
    y = xo.deform3d_cross_attn_forward(mod.state_dict(), sc["query"], feats2, sc["query_pos"], rp2,
                                       sc["metas"], syn.PC_RANGE, 8)
    y.square().sum().backward()
    assert H.rel_err(y, y_ref) <= 2e-6
    assert H.rel_err(rp2.grad, rp.grad) <= 1e-5
    for a, b in zip(feats2, feats):
        assert H.rel_err(a.grad, b.grad) <= 1e-5


def test_variant_c_module_batch2_reference_batch_order(ref):
    """B > 1 (quirk A.4-2): ``query.repeat(N,1,1)`` (deform3d_cross_attn.py:277) pairs image
    i = b*N+n with the attention logits of sample i % B.  The oracle reproduces exactly that with
    ``reference_batch_order=True`` and differs from it with per-sample logits -- which is why the
    product refuses B > 1 unless ``allow_batched=True``."""
    sc = H.scene(B=2, T=1, Q=60)
    torch.manual_seed(6)
    mod = ref.Deform3DCrossAttnCPU(num_cams=sc["N"], num_points=4, pc_range=syn.PC_RANGE).eval()
    syn.randomize_generators(mod)
    y_ref = _quiet(mod, sc["query"], None, sc["feats"], query_pos=sc["query_pos"], reference_points=sc["ref"],
                   img_metas=sc["metas"])
    args = (mod.state_dict(), sc["query"], sc["feats"], sc["query_pos"], sc["ref"], sc["metas"], syn.PC_RANGE, 8)
    y_quirk = xo.deform3d_cross_attn_forward(*args, reference_batch_order=True)
    y_sane = xo.deform3d_cross_attn_forward(*args, reference_batch_order=False)
    assert H.rel_err(y_quirk, y_ref) <= 2e-6
    assert H.rel_err(y_sane, y_ref) > 1e-3


def test_variant_v2_module(ref):
    sc = H.scene(B=1, T=1, Q=90)
    torch.manual_seed(5)
    mod = ref.Detr3DCrossAttenV2(num_cams=6, num_points=4, pc_range=syn.PC_RANGE).eval()
    syn.randomize_generators(mod)
    feats = [f.clone().requires_grad_(True) for f in sc["feats"]]
    rp = sc["ref"].clone().requires_grad_(True)
    y_ref = mod(sc["query"], None, feats, query_pos=sc["query_pos"], reference_points=rp, img_metas=sc["metas"])
    g = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(1))
    (y_ref * g).sum().backward()
    feats2 = [f.clone().requires_grad_(True) for f in sc["feats"]]
    rp2 = sc["ref"].clone().requires_grad_(True)
    y = xo.detr3d_cross_atten_v2_forward(mod.state_dict(), sc["query"], feats2, sc["query_pos"], rp2,
                                         sc["metas"], syn.PC_RANGE, 8)
    (y * g).sum().backward()
    assert H.rel_err(y, y_ref) <= 2e-6
    assert H.rel_err(rp2.grad, rp.grad) <= 1e-5
    for a, b in zip(feats2, feats):
        assert H.rel_err(a.grad, b.grad) <= 1e-5
    # the reference cannot run V2 unless num_points == num_levels (weights (L,P) x samples (P,L))
    bad = ref.Detr3DCrossAttenV2(num_cams=6, num_points=5, pc_range=syn.PC_RANGE).eval()
    with pytest.raises(RuntimeError):
        bad(sc["query"], None, sc["feats"], query_pos=sc["query_pos"], reference_points=sc["ref"],
            img_metas=sc["metas"])


def test_shipped_cpu_branch_is_broken_and_patch_is_one_token(ref):
    """deform3d_cross_attn.py:305-309: documents why the oracle needs the substitution."""
    sc = H.scene(B=1, T=1, Q=8)
    mod = ref.Deform3DCrossAttn(num_cams=6, num_points=4, pc_range=syn.PC_RANGE).eval()
    with pytest.raises(NameError):
        _quiet(mod, sc["query"], None, sc["feats"], query_pos=sc["query_pos"],
               reference_points=sc["ref"], img_metas=sc["metas"])


def test_product_modules_share_state_dict_keys_and_init_with_reference(ref):
    import graph_detr4d_b200 as g
    kw = dict(num_cams=12, num_points=4, pc_range=syn.PC_RANGE)
    r, m = ref.Deform3DCrossAttnCPU(**kw), g.Deform3DCrossAttn(**kw)
    assert list(r.state_dict().keys()) == list(m.state_dict().keys())
    for k, v in r.state_dict().items():
        assert m.state_dict()[k].shape == v.shape, k
    assert torch.equal(r.deform_sampling_offsets.bias, m.deform_sampling_offsets.bias)
    for name in ("cam_attention_weights", "attention_weights", "deform_sampling_offsets"):
        assert getattr(m, name).weight.abs().max() == 0
    kw = dict(num_cams=6, num_points=1, pc_range=syn.PC_RANGE)
    r, m = ref.Detr3DCrossAtten(**kw), g.Detr3DCrossAtten(**kw)
    assert list(r.state_dict().keys()) == list(m.state_dict().keys())
    m.load_state_dict(r.state_dict())
    import inspect
    assert list(inspect.signature(r.forward).parameters) == list(inspect.signature(m.forward).parameters)
