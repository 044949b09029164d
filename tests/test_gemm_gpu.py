"""GPU: the two fp32-accurate GEMM kernels behind include/gd4d_glue.h (row f2) against an fp64 product.
``gd4d_sgemm_small`` (register-tiled FFMA) and ``gd4d_gemm_tf32x3`` (tcgen05 3xTF32, fp32 accumulate in
tensor memory) must both stay within fp32-FMA-chain error of the exact result -- 2e-6 of max|C| at K <= 512,
i.e. the class of the cuBLAS sgemm they stand in for (measured 6e-7 / 6e-7 / 6e-7, profiles/r2_gemm_check.json)
-- for every operand layout the decoder's Linear sites need: x.W^T, dY.W, dY^T.X, batched, ragged sizes."""
import pytest
import torch

from graph_detr4d_b200 import gemm as G

pytestmark = pytest.mark.gpu

CASES = [
    # name, M, N, K, a_t, b_t, batch, bias, relu
    ("linear_fwd", 900, 256, 256, False, False, 0, True, False),
    ("ffn_relu", 900, 512, 256, False, False, 0, True, True),
    ("ffn_out_k512", 900, 256, 512, False, False, 0, True, False),
    ("generators_232", 900, 232, 256, False, False, 0, True, False),
    ("reg_branch_10", 900, 10, 256, False, False, 0, True, False),
    ("ragged", 70, 36, 40, False, False, 0, False, False),
    ("one_row_tile", 4, 8, 8, False, False, 0, True, True),
    ("dgrad", 900, 256, 512, False, True, 0, False, False),
    ("wgrad_batched", 256, 256, 900, True, True, 5, False, False),
    ("wide_value_proj", 900, 32, 256, False, False, 8, False, False),
    ("attn_pv", 900, 32, 900, False, True, 8, False, False),
]


@pytest.mark.parametrize("impl", ["simt", "tf32x3"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gemm_matches_fp64(case, impl):
    _, M, N, K, a_t, b_t, batch, bias, relu = case
    g = torch.Generator().manual_seed(M * 31 + N * 7 + K)
    sh = lambda r, c: ((batch, r, c) if batch else (r, c))
    a = torch.randn(sh(K, M) if a_t else sh(M, K), generator=g).cuda()
    b = torch.randn(sh(K, N) if b_t else sh(N, K), generator=g).cuda()
    bi = torch.randn(N, generator=g).cuda() if bias else None
    assert G.supported(a, b, a_t, b_t)
    A = a.transpose(-1, -2) if a_t else a
    B = b if b_t else b.transpose(-1, -2)
    ref = A.double() @ B.double()
    if bias:
        ref = ref + bi.double()
    if relu:
        ref = ref.relu()
    out = G.gemm(a, b, bi, relu, a_t, b_t, impl=impl)
    assert out.dtype == torch.float32 and tuple(out.shape) == tuple(ref.shape)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, err


def test_strided_views_and_out_argument():
    """Row-strided operands (a column block of a packed matrix) and writing into a strided ``out``."""
    g = torch.Generator().manual_seed(1)
    big_a = torch.randn(300, 512, generator=g).cuda()
    big_b = torch.randn(128, 768, generator=g).cuda()
    a, b = big_a[:, 128:384], big_b[:, 256:512]                    # (300,256) ld 512, (128,256) ld 768
    big_c = torch.zeros(300, 256, device="cuda")
    out = big_c[:, 64:192]
    for impl in ("simt", "tf32x3"):
        big_c.zero_()
        G.gemm(a, b, out=out, impl=impl)
        ref = a.double() @ b.double().t()
        assert float((out.double() - ref).abs().max() / ref.abs().max()) <= 2e-6
        assert float(big_c[:, :64].abs().max()) == 0.0 and float(big_c[:, 192:].abs().max()) == 0.0


def test_unsupported_shapes_are_reported_not_computed():
    a = torch.randn(10, 6, device="cuda")                          # K = 6: not a multiple of 4
    b = torch.randn(8, 6, device="cuda")
    assert not G.supported(a, b)
    with pytest.raises(ValueError):
        G.gemm(a, b)
    with pytest.raises(ValueError):
        G.gemm(torch.randn(8, 8), torch.randn(8, 8))                # CPU tensors: no fallback
