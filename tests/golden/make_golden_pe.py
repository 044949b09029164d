"""Golden vectors for row f4 (frustum position-embedding input): the UNMODIFIED reference method
Detr3DHeadPE.position_embeding executed in the build container with an identity position_encoder.
Run:  python tests/golden/make_golden_pe.py   ->  tests/golden/frustum_pe.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_pe_oracle import SHAPES, _masks, _reference  # noqa: E402

B, T, DEPTH_NUM = 2, 2, 8
masks = _masks(B, 6 * T, SHAPES)
xs, ms, _ = _reference(B, T, SHAPES, DEPTH_NUM, masks)
out = dict(B=B, T=T, depth_num=DEPTH_NUM, shapes=np.asarray(SHAPES))
for l, (x, m, mi) in enumerate(zip(xs, ms, masks)):
    out[f"x{l}"] = x.numpy()
    out[f"mask{l}"] = m.numpy()
    out[f"mask_in{l}"] = mi.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frustum_pe.npz"), **out)
print("wrote frustum_pe.npz", {k: getattr(v, "shape", v) for k, v in out.items()})
