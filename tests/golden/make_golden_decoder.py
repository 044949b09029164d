"""Golden vectors for row a9: the reference's OWN ``Detr3DTransformer`` +
``Detr3DTransformerDecoder`` classes (detr3d_transformer.py:45-225), executed unmodified through
oracle/ref_loader.py around the reference's own attention classes, at the real width (256 channels,
8 heads) on tiny maps so the product takes its default fused / wide path.  Build container only.
Run:  python tests/golden/make_golden_decoder.py

The fixture does NOT store the ~1.7 M weights or the feature maps: both are regenerated from seeds
(``build_case`` below, torch's CPU generator) and guarded by float64 checksums stored in the file -- a
torch whose generator or initialisers differ fails the checksum loudly instead of comparing garbage.
Stored: the decoder outputs, the refined reference points, and gradients of sum(states * gout) w.r.t.
the query embedding, the initial-reference Linear (layer 0 is the only layer whose reference points
carry gradient, :214), layer-0 generator / output_proj weights, and a
channel-strided (3::8) slice of every feature-map gradient.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graph_detr4d_b200 import synthetic as syn  # noqa: E402
from graph_detr4d_b200.decoder import Detr3DTransformer, Detr3DTransformerDecoder  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(8, 12), (4, 6), (2, 3), (1, 2)]
C, HEADS, Q, LAYERS, FFN = 256, 8, 64, 3, 512
CASES = {"A": dict(variant="A", T=1, B=1), "C": dict(variant="C", T=2, B=1), "A_B2": dict(variant="A", T=1, B=2)}


def build_case(name):
    """Seeded product-side model (CPU, parameters only), feature maps, img_metas and grad-out for a case.
    Shared by the generator, tests/test_decoder_oracle.py and tests/test_decoder_gpu.py."""
    cs = CASES[name]
    N = 6 * cs["T"]
    torch.manual_seed(41)
    if cs["variant"] == "A":
        cfg = dict(type="Detr3DCrossAtten", num_cams=N, num_points=1, pc_range=syn.PC_RANGE, dropout=0.0)
    else:
        cfg = dict(type="Deform3DCrossAttn", num_cams=N, num_points=4, pc_range=syn.PC_RANGE, dropout=0.0)
    dec = Detr3DTransformerDecoder(cfg, num_layers=LAYERS, embed_dims=C, num_heads=HEADS,
                                   feedforward_channels=FFN, dropout=0.0)
    model = Detr3DTransformer(dec, num_query=Q)
    for i, layer in enumerate(dec.layers):
        syn.randomize_generators(layer.attentions[1], seed=50 + i)
    feats = [f.to(torch.bfloat16).float() for f in syn.make_feats(cs["B"], N, C, SHAPES, seed=43)]
    metas = syn.make_img_metas(cs["B"], cs["T"])
    gout = torch.randn(LAYERS, Q, cs["B"], C, generator=torch.Generator().manual_seed(44))
    return model.eval(), feats, metas, gout, cs


def checksum(tensors):
    return np.array([float(t.detach().double().abs().sum()) for t in tensors])


def run_reference(name):
    """Execute the reference classes on the case; returns (dict of outputs, product-side model)."""
    from oracle import decoder_oracle, ref_loader
    ref = ref_loader.load()
    model, feats, metas, gout, cs = build_case(name)
    N = 6 * cs["T"]
    rdec = decoder_oracle.build_reference_decoder(ref, cs["variant"], N, LAYERS, embed_dims=C, num_heads=HEADS,
                                                  num_points=4, feedforward_channels=FFN, pc_range=syn.PC_RANGE)
    rdec.embed_dims = C
    rmodel = ref.Detr3DTransformer(num_cams=N, decoder=rdec).eval()
    sd = model.state_dict()
    own = {k: v for k, v in sd.items() if not k.startswith(("query_embedding.", "reg_branches."))}
    missing, unexpected = rmodel.load_state_dict(own, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    import copy
    embed = model.query_embedding.weight.detach().clone().requires_grad_(True)
    branches = copy.deepcopy(model.reg_branches)
    feats = [f.clone().requires_grad_(True) for f in feats]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        st, r0, refs = rmodel(feats, embed, reg_branches=branches, img_metas=metas)
    (st * gout).sum().backward()
    attn0 = rmodel.decoder.layers[0].attentions[1]
    out = dict(states=st.detach().numpy(), init_ref=r0.detach().numpy(), refs=refs.detach().numpy(),
               grad_embed=embed.grad.numpy(),
               grad_refpoint_w=rmodel.reference_points.weight.grad.numpy(),
               grad_attnw0=attn0.attention_weights.weight.grad.numpy(),
               grad_outproj0=attn0.output_proj.weight.grad.numpy())
    assert all(p.grad is None for p in branches.parameters())      # refined points are detached (:214)
    for i, f in enumerate(feats):
        out[f"grad_feat{i}_s8"] = f.grad[:, :, 3::8].contiguous().numpy()
    out["sum_params"] = checksum(sd.values())
    out["sum_feats"] = checksum(feats)
    return out, model


def main():
    import warnings
    warnings.filterwarnings("ignore")
    for name in CASES:
        out, _ = run_reference(name)
        path = os.path.join(OUT, f"decoder_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KB", "states absmax", float(np.abs(out["states"]).max()))


if __name__ == "__main__":
    main()
