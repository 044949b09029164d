"""Row a9 pinned to the EXECUTED reference (CPU).

  * in the build container (reference tree present): the reference's own ``Detr3DTransformer`` +
    ``Detr3DTransformerDecoder`` classes are executed again (tests/golden/make_golden_decoder.run_reference)
    and must reproduce the committed fixture -- the fixture is the reference's output, not ours;
  * everywhere: graph_detr4d_b200.decoder's layer loop / refinement (host logic, torch ops on CPU tensors)
    built around the CPU oracle's attention must reproduce the fixture: loop order, post-norm placement,
    logit-space refinement of x,y and z (reg columns 0,1 and 4), detach after every layer, layer-0
    reference gradient (detr3d_transformer.py:192-214, :128-147).
The GPU run of the same comparison is tests/test_decoder_gpu.py::test_decoder_matches_reference_golden."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle.modules_port import build_oracle_attention
from tests import helpers as H
from tests.golden import make_golden_decoder as mg

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_decoder_golden(name):
    gd = np.load(os.path.join(GOLD_DIR, f"decoder_{name}.npz"), allow_pickle=False)
    return {k: torch.as_tensor(gd[k]) for k in gd.files}


def check_inputs_regenerated(gd, model, feats):
    """The fixture holds no weights / maps: the seeded regeneration must be the one the fixture saw."""
    assert np.array_equal(mg.checksum(model.state_dict().values()), gd["sum_params"].numpy()), \
        "seeded weights differ from the ones the golden was generated with (torch RNG / init changed)"
    assert np.array_equal(mg.checksum(feats), gd["sum_feats"].numpy())


def compare_with_golden(gd, st, r0, refs, embed_grad, refpoint_w_grad, attn0, feat_grads, tol, gtol):
    assert H.rel_err(st, gd["states"]) <= tol
    assert H.rel_err(r0, gd["init_ref"]) <= tol
    assert H.rel_err(refs, gd["refs"]) <= tol
    assert H.rel_err(embed_grad, gd["grad_embed"]) <= gtol
    assert H.rel_err(refpoint_w_grad, gd["grad_refpoint_w"]) <= gtol          # only layer 0 feeds it (:214)
    assert H.rel_err(attn0.attention_weights.weight.grad.cpu(), gd["grad_attnw0"]) <= gtol
    assert H.rel_err(attn0.output_proj.weight.grad.cpu(), gd["grad_outproj0"]) <= gtol
    for i, g in enumerate(feat_grads):
        assert H.rel_err(g[:, :, 3::8], gd[f"grad_feat{i}_s8"]) <= gtol


@pytest.mark.reference
@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", list(mg.CASES))
def test_fixture_is_the_executed_reference(name):
    out, _ = mg.run_reference(name)
    gd = load_decoder_golden(name)
    assert set(out) == set(gd)
    for k, v in out.items():
        assert H.rel_err(torch.as_tensor(v), gd[k]) <= 1e-6, k  # same code, same BLAS: only thread-count reorderings


@pytest.mark.parametrize("name", list(mg.CASES))
def test_decoder_loop_on_cpu_matches_reference_golden(name):
    from graph_detr4d_b200.decoder import Detr3DTransformer, Detr3DTransformerDecoder
    gd = load_decoder_golden(name)
    seeded, feats, metas, gout, cs = mg.build_case(name)
    check_inputs_regenerated(gd, seeded, feats)
    N = 6 * cs["T"]
    cfg = dict(type="Detr3DCrossAtten" if cs["variant"] == "A" else "Deform3DCrossAttn", num_cams=N,
               num_points=1 if cs["variant"] == "A" else 4, pc_range=mg.syn.PC_RANGE, dropout=0.0)
    dec = Detr3DTransformerDecoder(cfg, num_layers=mg.LAYERS, embed_dims=mg.C, num_heads=mg.HEADS,
                                   feedforward_channels=mg.FFN, dropout=0.0, cross_attn_factory=build_oracle_attention)
    model = Detr3DTransformer(dec, num_query=mg.Q).eval()
    model.load_state_dict(seeded.state_dict(), strict=True)
    feats = [f.clone().requires_grad_(True) for f in feats]
    st, r0, refs = model(feats, metas, cs["B"])
    (st * gout).sum().backward()
    assert all(p.grad is None for p in model.reg_branches.parameters())      # refined points are detached
    compare_with_golden(gd, st.detach(), r0.detach(), refs.detach(), model.query_embedding.weight.grad,
                        model.reference_points.weight.grad, model.decoder.layers[0].attentions[1],
                        [f.grad for f in feats], tol=2e-5, gtol=2e-4)
