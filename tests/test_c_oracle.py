"""The independent plain-C restatement (oracle/c/xview_ref.c) agrees with the torch oracle:
mask bit-for-bit, values to 1e-5 (different summation order)."""
import numpy as np
import pytest
import torch

from graph_detr4d_b200 import synthetic as syn
from oracle import c_ref, xview_oracle as xo
from tests import helpers as H


@pytest.mark.parametrize("B,T,P", [(1, 1, 1), (2, 2, 2)])
def test_c_restatement_mode_a(B, T, P):
    sc = H.scene(B=B, T=T, Q=60, C=64, shapes=[(12, 20), (6, 10), (3, 5), (2, 3)])
    logits = H.rand_inputs_a(sc, P=P)
    out_t, mask_t = xo.xview_a_core(sc["feats"], sc["ref"], logits, sc["l2i"], syn.PC_RANGE, 900, 1600)
    out_c, mask_c = c_ref.forward(0, [f.numpy() for f in sc["feats"]], sc["ref"].numpy(), logits.numpy(),
                                  sc["l2i"].numpy(), syn.PC_RANGE, 900.0, 1600.0, 2, P)
    assert np.array_equal(mask_c.astype(bool), mask_t.numpy())
    assert mask_t.any()
    assert H.rel_err(torch.from_numpy(out_c), out_t) <= 1e-5


@pytest.mark.parametrize("B,T,P", [(1, 2, 4), (2, 1, 3)])
def test_c_restatement_mode_c(B, T, P):
    sc = H.scene(B=B, T=T, Q=40, C=64, shapes=[(12, 20), (6, 10), (3, 5), (2, 3)])
    logits, offsets, cam = H.rand_inputs_c(sc, Hh=2, P=P)
    out_t, mask_t = xo.xview_c_core(sc["feats"], sc["ref"], offsets, logits, cam, sc["l2i"], syn.PC_RANGE,
                                    900, 1600, 2)
    out_c, mask_c = c_ref.forward(1, [f.numpy() for f in sc["feats"]], sc["ref"].numpy(), logits.numpy(),
                                  sc["l2i"].numpy(), syn.PC_RANGE, 900.0, 1600.0, 2, P,
                                  offsets=offsets.numpy(), cam_logits=cam.numpy())
    assert np.array_equal(mask_c.astype(bool), mask_t[:, :, :, :, 0, :].numpy())
    assert mask_t.any()
    assert H.rel_err(torch.from_numpy(out_c), out_t) <= 1e-5


def test_c_restatement_frustum_pe():
    """Second, independent restatement (plain C) of the frustum position-embedding input vs the torch oracle."""
    import numpy as np
    import torch
    from graph_detr4d_b200 import synthetic as syn
    from oracle import c_ref, pe_oracle
    shapes, D, T, B = [(6, 10), (3, 5)], 8, 2, 2
    metas = syn.make_img_metas(B, T)
    g = torch.Generator().manual_seed(3)
    masks = [torch.rand(B, 6 * T, H, W, generator=g) > 0.8 for H, W in shapes]
    xo, mo = pe_oracle.frustum_pe_input(shapes, metas, D, 1, syn.PC_RANGE, masks)
    i2l = pe_oracle.img2lidar_fp32(metas).numpy()
    bin_size = np.float32((syn.PC_RANGE[3] - 1) / (D * (1 + D)))
    for (H, W), x, m, mi in zip(shapes, xo, mo, masks):
        out, mask = c_ref.pe_frustum(i2l, mi.numpy(), H, W, D, 928.0, 1600.0, 1.0, float(bin_size), syn.PC_RANGE)
        assert np.array_equal(mask.reshape(m.shape).astype(bool), m.numpy())            # mask bit-exact
        assert float(np.abs(out - x.numpy()).max()) <= 1e-5 * float(x.abs().max())     # libm logf ulp only


def test_c_restatement_match_cost():
    import numpy as np
    from oracle import assign_oracle as ao, c_ref
    from tests.test_assign_oracle import make_case
    bbox, cls, gt, labels = make_case(200, 23, seed=4)
    want = ao.match_cost(bbox, cls, gt, labels).numpy()
    got = c_ref.match_cost(cls.numpy(), bbox.numpy(), gt.numpy(), labels.numpy())
    assert float(np.abs(got - want).max()) <= 1e-5 * float(np.abs(want).max())
