"""Host logic of DeferredWgrad.flush() (pure torch, runs on CPU): batched weight / bias / LayerNorm
gradients equal the per-layer formulas, packed in_proj parameters are assembled from their row
slices (batched cat across layers), and the buffers / covered bookkeeping that the N > 1 step
all-reduces in place really covers every assigned gradient."""
import torch

from graph_detr4d_b200.glue import DeferredWgrad


def _fill(q, layers=4, M=50, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    want, params = {}, []
    for _ in range(layers):
        inw, inb = torch.nn.Parameter(torch.zeros(24, 8)), torch.nn.Parameter(torch.zeros(24))
        ow, ob = torch.nn.Parameter(torch.zeros(8, 8)), torch.nn.Parameter(torch.zeros(8))
        fw, fb = torch.nn.Parameter(torch.zeros(16, 8)), torch.nn.Parameter(torch.zeros(16))
        lw, lb = torch.nn.Parameter(torch.ones(8)), torch.nn.Parameter(torch.zeros(8))
        x1, gqk, x2, gv, x3, go, x4, gf = r(M, 8), r(M, 16), r(M, 8), r(M, 8), r(M, 8), r(M, 8), r(M, 8), r(M, 16)
        q.items += [(ow, ob, 0, 8, go, x3), (inw, inb, 16, 24, gv, x2), (inw, inb, 0, 16, gqk, x1), (fw, fb, 0, 16, gf, x4)]
        gl, xl = r(M, 8), r(M, 8)
        mean, var = xl.mean(1, keepdim=True), xl.var(1, unbiased=False, keepdim=True)
        rstd = (var + 1e-5).rsqrt()
        q.ln_items.append((lw, lb, gl, xl, mean, rstd))
        want[inw] = torch.cat([gqk.t() @ x1, gv.t() @ x2]); want[inb] = torch.cat([gqk.sum(0), gv.sum(0)])
        want[ow] = go.t() @ x3; want[ob] = go.sum(0)
        want[fw] = gf.t() @ x4; want[fb] = gf.sum(0)
        want[lw] = (gl * (xl - mean) * rstd).sum(0); want[lb] = gl.sum(0)
        params += [inw, inb, ow, ob, fw, fb, lw, lb]
    return want, params


def test_flush_matches_per_layer_formulas_and_bookkeeping():
    q = DeferredWgrad()
    want, params = _fill(q)
    q.flush()
    assert not q.items and not q.ln_items
    for p in params:
        assert p.grad is not None and torch.allclose(p.grad, want[p], atol=1e-4), tuple(p.shape)
    # every parameter is covered, and its gradient is a view into one of the recorded buffers
    assert q.covered == {id(p) for p in params}
    spans = [(b.untyped_storage().data_ptr(), b.data_ptr(), b.data_ptr() + b.numel() * 4) for b in q.buffers]
    for p in params:
        lo = p.grad.data_ptr()
        assert any(a <= lo and lo + p.grad.numel() * 4 <= hi for _, a, hi in spans), tuple(p.shape)
    # in-place all-reduce of the buffers (here: scaling) must reach every parameter's gradient
    for b in q.buffers:
        b.mul_(0.5)
    for p in params:
        assert torch.allclose(p.grad, want[p] * 0.5, atol=1e-4)
    # a handful of buffers, not one per parameter
    assert len(q.buffers) <= 12 < len(params)


def test_second_contribution_accumulates_and_uncovers():
    q = DeferredWgrad()
    w, b = torch.nn.Parameter(torch.zeros(8, 8)), torch.nn.Parameter(torch.zeros(8))
    w.grad, b.grad = torch.ones(8, 8), torch.ones(8)                  # e.g. a directly produced gradient
    g, x = torch.randn(20, 8), torch.randn(20, 8)
    q.items.append((w, b, 0, 8, g, x))
    q.flush()
    assert torch.allclose(w.grad, torch.ones(8, 8) + g.t() @ x, atol=1e-5)
    assert id(w) not in q.covered and id(b) not in q.covered          # must be reduced on its own
