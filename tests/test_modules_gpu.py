"""GPU: the drop-in modules vs the committed golden vectors (reference outputs) and
vs the oracle's whole-module restatement."""
import pytest
import torch

import graph_detr4d_b200 as g
from graph_detr4d_b200 import synthetic as syn
from oracle import xview_oracle as xo
from tests import helpers as H
from tests.test_oracle_golden import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant,cls", [("A", "Detr3DCrossAtten"), ("C", "Deform3DCrossAttn"),
                                         ("V2", "Detr3DCrossAttenV2"), ("C256", "Deform3DCrossAttn")])
def test_module_matches_reference_golden(variant, cls):
    """Outputs and gradients frozen from the UNMODIFIED reference classes (tests/golden/make_golden.py).
    "C256" has 1 KB pixel rows, so the module runs its default wide (gather-then-project) kernels."""
    gd = load_golden(variant)
    N = 6 * gd["T"]
    m = getattr(g, cls)(embed_dims=gd["C"], num_heads=gd["heads"], num_levels=4,
                        num_points=1 if variant == "A" else 4, num_cams=N, pc_range=syn.PC_RANGE).cuda().eval()
    m.load_state_dict(gd["sd"], strict=True)
    feats = [f.cuda().requires_grad_(True) for f in gd["feats"]]
    q = gd["query"].cuda().requires_grad_(True)
    rp = gd["ref"].cuda().requires_grad_(True)
    g.clear_caches()
    y = m(q, None, feats, query_pos=gd["query_pos"].cuda(), reference_points=rp, img_metas=gd["metas"])
    (y * gd["gout"].cuda()).sum().backward()
    assert tuple(y.shape) == tuple(gd["out"].shape)
    if variant == "C256":
        assert m._use_wide(g.ops.pack_features([f.detach() for f in feats]))
    assert H.rel_err(y.detach().cpu(), gd["out"]) <= 1e-5
    assert H.rel_err(q.grad.cpu(), gd["grad_query"]) <= 2e-4
    assert H.rel_err(rp.grad.cpu(), gd["grad_ref"]) <= 2e-4
    for f, gref in zip(feats, gd["grad_feats"]):
        assert H.rel_err(f.grad.cpu(), gref) <= 2e-4


@pytest.mark.parametrize("variant", ["A", "C"])
def test_module_256ch_vs_oracle_and_layer_sharing(variant):
    T = 1 if variant == "A" else 2
    sc = H.scene(B=1, T=T, Q=200)
    torch.manual_seed(7)
    if variant == "A":
        m = g.Detr3DCrossAtten(num_cams=sc["N"], num_points=1, pc_range=syn.PC_RANGE)
    else:
        m = g.Deform3DCrossAttn(num_cams=sc["N"], num_points=4, pc_range=syn.PC_RANGE)
    syn.randomize_generators(m)
    m = m.cuda().eval()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    if variant == "A":
        y_ref = xo.detr3d_cross_atten_forward(sd, sc["query"], sc["feats"], sc["query_pos"], sc["ref"],
                                              sc["metas"], syn.PC_RANGE)
    else:
        y_ref = xo.deform3d_cross_attn_forward(sd, sc["query"], sc["feats"], sc["query_pos"], sc["ref"],
                                               sc["metas"], syn.PC_RANGE, 8)
    feats = [f.cuda() for f in sc["feats"]]
    g.clear_caches()
    with torch.no_grad():
        y1 = m(sc["query"].cuda(), None, feats, query_pos=sc["query_pos"].cuda(),
               reference_points=sc["ref"].cuda(), img_metas=sc["metas"])
        packed_first = g.modules._PACK_CACHE._packed
        y2 = m(sc["query"].cuda(), None, feats, query_pos=sc["query_pos"].cuda(),
               reference_points=sc["ref"].cuda(), img_metas=sc["metas"])
    assert g.modules._PACK_CACHE._packed is packed_first        # second "layer" reused the packed maps
    assert torch.equal(y1, y2)
    # value_proj runs in cuBLAS fp32 vs the CPU oracle's GEMM: allow GEMM-order noise
    assert H.rel_err(y1.cpu(), y_ref) <= (1e-5 if variant == "A" else 5e-5)
    feats[0].add_(1.0)                                          # in-place change must invalidate the cache
    with torch.no_grad():
        y3 = m(sc["query"].cuda(), None, feats, query_pos=sc["query_pos"].cuda(),
               reference_points=sc["ref"].cuda(), img_metas=sc["metas"])
    assert not torch.equal(y1, y3)


def test_module_bf16_features():
    sc = H.scene(B=1, T=2, Q=150)
    torch.manual_seed(7)
    m = g.Deform3DCrossAttn(num_cams=12, num_points=4, pc_range=syn.PC_RANGE, feature_dtype="bf16")
    syn.randomize_generators(m)
    m = m.cuda().eval()
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    y_ref = xo.deform3d_cross_attn_forward(sd, sc["query"], sc["feats"], sc["query_pos"], sc["ref"],
                                           sc["metas"], syn.PC_RANGE, 8)
    g.clear_caches()
    with torch.no_grad():
        y = m(sc["query"].cuda(), None, [f.cuda() for f in sc["feats"]], query_pos=sc["query_pos"].cuda(),
              reference_points=sc["ref"].cuda(), img_metas=sc["metas"])
    # bf16 features AND a bf16 value_proj GEMM: stated tolerance 2e-2 of max|ref|
    assert H.rel_err(y.cpu(), y_ref) <= 2e-2
