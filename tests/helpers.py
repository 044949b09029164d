"""Shared builders for the parity tests (seeded synthetic scenes)."""
import numpy as np
import torch

from graph_detr4d_b200 import synthetic as syn

SMALL_SHAPES = [(29, 50), (15, 25), (8, 13), (4, 7)]       # 232x400 padded input, strides 8..64
FULL_SHAPES = syn.LEVEL_SHAPES_928x1600


def scene(B=1, T=1, Q=128, shapes=SMALL_SHAPES, C=256, seed=0):
    N = 6 * T
    feats = syn.make_feats(B, N, C, shapes, seed=seed)
    query, query_pos, ref = syn.make_queries(B, Q, C, seed=seed + 1)
    metas = syn.make_img_metas(B, T)
    l2i = torch.as_tensor(np.asarray([m["lidar2img"] for m in metas]).astype(np.float32))
    return dict(B=B, N=N, Q=Q, C=C, feats=feats, query=query, query_pos=query_pos, ref=ref,
                metas=metas, l2i=l2i, img_h=900.0, img_w=1600.0)


def rand_inputs_a(sc, P=1, seed=5, std=1.0):
    g = torch.Generator().manual_seed(seed)
    L = len(sc["feats"])
    return torch.randn(sc["B"], sc["Q"], sc["N"] * P * L, generator=g) * std


def rand_inputs_c(sc, Hh=8, P=4, seed=6, off_std=2.0):
    g = torch.Generator().manual_seed(seed)
    L = len(sc["feats"])
    B, Q, N = sc["B"], sc["Q"], sc["N"]
    logits = torch.randn(B, Q, Hh * L * P, generator=g)
    offsets = torch.randn(B, Q, Hh * P * 3, generator=g) * off_std
    cam = torch.randn(B, Q, N, generator=g)
    return logits, offsets, cam


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
