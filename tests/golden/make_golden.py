"""Generate golden vectors by executing the UNMODIFIED reference (build container
only: needs /root/reference).  Run:  python tests/golden/make_golden.py

Each .npz holds the inputs (features stored as bf16 bit patterns: exactly
representable, half the bytes), the reference module's state_dict, its eval-mode
forward output and the gradients of sum(out * gout) w.r.t. query, reference
points and every feature level.  Small on purpose (embed_dims=64 -> 2 heads of
32 channels) so the fixtures stay a few hundred KB.

Variant "C256" is Deform3DCrossAttn at the REAL width (embed_dims=256, 8 heads, 6 cameras) on tiny
maps (8x12 .. 1x2): with 1 KB pixel rows the product takes its default "wide" (gather-then-project)
kernels -- the ones bench.py times -- so a reference-held fixture pins that path too.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graph_detr4d_b200 import synthetic as syn  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(12, 20), (6, 10), (3, 5), (2, 3)]
C, HEADS, Q = 64, 2, 48


def bf16_bits(t):
    return t.to(torch.bfloat16).view(torch.int16).numpy()


def run(variant):
    global C, HEADS, Q, SHAPES
    ref = ref_loader.load()
    T = 2 if variant == "C" else 1
    N = 6 * T
    if variant == "C256":
        C, HEADS, Q, SHAPES = 256, 8, 64, [(8, 12), (4, 6), (2, 3), (1, 2)]
    else:
        C, HEADS, Q, SHAPES = 64, 2, 48, [(12, 20), (6, 10), (3, 5), (2, 3)]
    feats = [f.to(torch.bfloat16).float() for f in syn.make_feats(1, N, C, SHAPES, seed=11)]
    query, query_pos, rp = syn.make_queries(1, Q, C, seed=12)
    metas = syn.make_img_metas(1, T)
    torch.manual_seed(13)
    if variant == "A":
        mod = ref.Detr3DCrossAtten(embed_dims=C, num_heads=HEADS, num_levels=4, num_points=1, num_cams=N,
                                   pc_range=syn.PC_RANGE, dropout=0.1)
    elif variant == "V2":
        mod = ref.Detr3DCrossAttenV2(embed_dims=C, num_heads=HEADS, num_levels=4, num_points=4, num_cams=N,
                                     pc_range=syn.PC_RANGE, dropout=0.1)
    else:
        mod = ref.Deform3DCrossAttnCPU(embed_dims=C, num_heads=HEADS, num_levels=4, num_points=4,
                                       num_cams=N, pc_range=syn.PC_RANGE, dropout=0.1)
    mod.eval()
    syn.randomize_generators(mod, std=0.05, seed=14)
    feats = [f.requires_grad_(True) for f in feats]
    query = query.requires_grad_(True)
    rp = rp.requires_grad_(True)
    with contextlib.redirect_stdout(io.StringIO()):
        out = mod(query, None, feats, query_pos=query_pos, reference_points=rp, img_metas=metas)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(15))
    (out * gout).sum().backward()
    data = dict(
        variant=np.array(variant), num_frames=np.array(T), shapes=np.array(SHAPES),
        embed_dims=np.array(C), num_heads=np.array(HEADS),
        query=query.detach().numpy(), query_pos=query_pos.numpy(), ref=rp.detach().numpy(),
        gout=gout.numpy(), out=out.detach().numpy(),
        grad_query=query.grad.numpy(), grad_ref=rp.grad.numpy(),
        lidar2img=np.asarray(metas[0]["lidar2img"]),
    )
    for i, f in enumerate(feats):
        data[f"feat{i}_bf16"] = bf16_bits(f.detach())
        data[f"grad_feat{i}"] = f.grad.numpy()
    for k, v in mod.state_dict().items():
        data["sd." + k] = v.numpy()
    if variant == "A":
        _, _, mask = ref.feature_sampling([f.detach() for f in feats], rp.detach(), syn.PC_RANGE, metas)
        data["mask"] = mask.numpy()
    path = os.path.join(OUT, f"module_{variant}.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path) // 1024, "KB", "out max", float(out.abs().max()))


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore")
    only = sys.argv[1:] or ["A", "C", "V2", "C256"]
    for v in only:
        run(v)
