"""Seeded synthetic nuScenes-shaped inputs for the cross-view sampling path.

The construction is the one SURVEY.md section 8(d) / Appendix B fixes so that
every test, the bench and the golden-vector generator see the same scenes:
a deterministic 6-camera rig (yaws 0,-55,55,180,110,-110 degrees), past frame
``t`` shifted -4 m * t along x and appended on the camera axis (the reference
concatenates temporal frames as extra cameras:
projects/mmdet3d_plugin/datasets/pipelines/loading.py:120-183), FPN level sizes
of a 928x1600 padded input at strides 8/16/32/64, ``img_shape`` = the unpadded
(900, 1600, 3) that the reference normalises by (detr3d_transformer.py:419-420).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]          # projects/configs/detr3d/detr3d_res50.py:10
IMG_SHAPE = (900, 1600, 3)                                # unpadded (transform_3d.py:56-58)
PAD_SHAPE = (928, 1600, 3)
LEVEL_SHAPES_928x1600 = [(116, 200), (58, 100), (29, 50), (15, 25)]
_YAWS_DEG = [0.0, -55.0, 55.0, 180.0, 110.0, -110.0]
_FX = [1266.0, 1260.0, 1257.0, 809.0, 1256.0, 1259.0]
_CX, _CY = 803.3, 491.7


def level_shapes(pad_hw: Tuple[int, int] = (928, 1600), strides=(8, 16, 32, 64)):
    h, w = pad_hw
    return [(math.ceil(h / s), math.ceil(w / s)) for s in strides]


def make_lidar2img(num_frames: int = 1, dtype=np.float64) -> np.ndarray:
    """(6*T, 4, 4) lidar->image matrices of the synthetic rig (Appendix B)."""
    mats = []
    for t in range(num_frames):
        for yaw_deg, fx in zip(_YAWS_DEG, _FX):
            psi = math.radians(yaw_deg)
            fwd = np.array([math.cos(psi), math.sin(psi), 0.0])
            right = np.array([math.sin(psi), -math.cos(psi), 0.0])
            down = np.array([0.0, 0.0, -1.0])
            R = np.stack([right, down, fwd])
            c = np.array([1.5 * math.cos(psi), 0.5 * math.sin(psi), -0.3]) + np.array([-4.0 * t, 0.0, 0.0])
            E = np.eye(4)
            E[:3, :3] = R
            E[:3, 3] = -R @ c
            K = np.array([[fx, 0, _CX, 0], [0, fx, _CY, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
            mats.append(K @ E)
    return np.asarray(mats, dtype=dtype)


def make_img_metas(batch: int, num_frames: int = 1, img_shape=IMG_SHAPE, pad_shape=PAD_SHAPE):
    """img_metas as the reference's dataset pipeline produces them."""
    l2i = make_lidar2img(num_frames)
    n = l2i.shape[0]
    metas = []
    for b in range(batch):
        mats = [l2i[i].copy() for i in range(n)]
        if b > 0:  # make samples differ: small extra ego shift per sample
            shift = np.eye(4)
            shift[0, 3] = 0.7 * b
            shift[1, 3] = -0.4 * b
            mats = [m @ shift for m in mats]
        metas.append(dict(lidar2img=mats,
                          img_shape=[tuple(img_shape)] * n,
                          pad_shape=[tuple(pad_shape)] * n))
    return metas


def make_feats(batch: int, num_cams: int, channels: int = 256,
               shapes: Sequence[Tuple[int, int]] = LEVEL_SHAPES_928x1600,
               seed: int = 0, device="cpu", dtype=torch.float32) -> List[torch.Tensor]:
    """4 FPN levels, NCHW per camera: list of (B, N, C, H_l, W_l) (detr3d.py:39-66)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    feats = []
    for (h, w) in shapes:
        f = torch.randn(batch, num_cams, channels, h, w, generator=g, dtype=torch.float32)
        feats.append(f.to(device=device, dtype=dtype))
    return feats


def make_queries(batch: int, num_query: int, channels: int = 256, seed: int = 1, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    query = torch.randn(num_query, batch, channels, generator=g)
    query_pos = torch.randn(num_query, batch, channels, generator=g)
    ref = torch.rand(batch, num_query, 3, generator=g)
    return query.to(device), query_pos.to(device), ref.to(device)


def randomize_generators(module: torch.nn.Module, std: float = 0.05, seed: int = 2):
    """Reference init leaves the weight/offset generators at zero; give them
    N(0, std) weights so parity checks see non-trivial weights."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    for name in ("attention_weights", "cam_attention_weights", "deform_sampling_offsets",
                 "sampling_offsets"):
        lin = getattr(module, name, None)
        if lin is not None:
            with torch.no_grad():
                lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * std)
                lin.bias.add_(torch.randn(lin.bias.shape, generator=g) * std)
    return module
