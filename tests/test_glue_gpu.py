"""Fused per-layer glue kernels (include/gd4d_glue.h) against the reference's op-by-op torch
formulation evaluated in fp32 on the CPU: inverse_sigmoid (both clamp variants, ties and
out-of-range inputs), the logit-space reference-point refinement, and residual-sum +
LayerNorm (+ReLU) forward/backward including the deferred gamma/beta reduction."""
import pytest
import torch
import torch.nn.functional as F

from graph_detr4d_b200 import fused, modules
from graph_detr4d_b200.glue import DeferredWgrad

pytestmark = pytest.mark.gpu

TOL = 1e-5          # relative to max |ref|, fp32


def _rel(a, b):
    return float((a.cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _edge_points():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(900, 3, generator=g)
    x.view(-1)[:12] = torch.tensor([0.0, 1.0, 1e-5, 1 - 1e-5, 5e-6, 1 - 5e-6, -0.3, 1.7, 0.5, 1e-5 * 0.999,
                                    0.25, 0.75])
    return x


@pytest.mark.parametrize("clamp_max", [False, True])
def test_inverse_sigmoid_forward_backward(clamp_max):
    x = _edge_points()
    xo = x.clone().requires_grad_(True)
    yo = modules.inverse_sigmoid(xo, clamp_max=clamp_max)              # reference formulation, CPU
    g = torch.randn(x.shape, generator=torch.Generator().manual_seed(1))
    yo.backward(g)
    xg = x.cuda().requires_grad_(True)
    yg = fused.inverse_sigmoid(xg, 1e-5, clamp_max)
    yg.backward(g.cuda())
    assert _rel(yg.detach(), yo.detach()) <= TOL
    assert _rel(xg.grad, xo.grad) <= TOL
    # clamped-out inputs get exactly zero gradient, like torch.clamp
    assert float(xg.grad.view(-1)[6]) == 0.0 and float(xg.grad.view(-1)[7]) == 0.0


def test_reference_point_refinement_matches_reference_lines():
    """detr3d_transformer.py:201-214 restated op by op."""
    g = torch.Generator().manual_seed(2)
    ref = torch.rand(2, 900, 3, generator=g)
    ref[0, 0] = torch.tensor([0.0, 1.0, 1e-6])
    reg = torch.randn(2, 900, 10, generator=g)
    new = torch.zeros_like(ref)
    new[..., :2] = reg[..., :2] + modules.inverse_sigmoid(ref[..., :2])
    new[..., 2:3] = reg[..., 4:5] + modules.inverse_sigmoid(ref[..., 2:3])
    want = new.sigmoid()
    got = fused.ref_update(reg.cuda().requires_grad_(True), ref.cuda())
    assert not got.requires_grad
    assert float((got.cpu() - want).abs().max()) <= 1e-6


@pytest.mark.parametrize("C", [128, 256, 512])
@pytest.mark.parametrize("nres,relu", [(0, True), (0, False), (1, False), (2, False), (2, True)])
def test_add_layernorm_matches_torch(C, nres, relu):
    g = torch.Generator().manual_seed(C + nres)
    rows = 901                                                          # not a multiple of the CTA's rows
    xs = [torch.randn(rows, 1, C, generator=g) for _ in range(1 + nres)]
    ln = torch.nn.LayerNorm(C)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(C, generator=g)); ln.bias.copy_(torch.randn(C, generator=g) * 0.3)
    gy = torch.randn(rows, 1, C, generator=g)

    xo = [t.clone().requires_grad_(True) for t in xs]
    yo = F.layer_norm(sum(xo), (C,), ln.weight, ln.bias, ln.eps)
    if relu:
        yo = yo.relu()
    yo.backward(gy)
    want_w, want_b = ln.weight.grad.clone(), ln.bias.grad.clone()

    lng = torch.nn.LayerNorm(C).cuda()
    lng.load_state_dict(ln.state_dict())
    xg = [t.cuda().requires_grad_(True) for t in xs]
    yg = fused.add_layernorm(xg[0], lng, *xg[1:], relu=relu)
    yg.backward(gy.cuda())
    assert _rel(yg.detach(), yo.detach()) <= TOL
    for a, b in zip(xg, xo):
        assert _rel(a.grad, b.grad) <= 5e-5
    assert _rel(lng.weight.grad, want_w) <= 5e-5 and _rel(lng.bias.grad, want_b) <= 5e-5

    # deferred (batched) gamma/beta gradients give the same numbers
    lng.zero_grad(set_to_none=True)
    xg = [t.cuda().requires_grad_(True) for t in xs]
    with DeferredWgrad() as wq:
        fused.add_layernorm(xg[0], lng, *xg[1:], relu=relu).backward(gy.cuda())
        assert lng.weight.grad is None
        wq.flush()
    assert _rel(lng.weight.grad, want_w) <= 5e-5 and _rel(lng.bias.grad, want_b) <= 5e-5


def test_linear_relu_and_deferred_bias_match_torch():
    """GEMM + one bias(+ReLU) launch, and GEMM-without-epilogue + bias folded into add_layernorm."""
    from graph_detr4d_b200 import glue
    g = torch.Generator().manual_seed(11)
    lin, ln = torch.nn.Linear(256, 512), torch.nn.LayerNorm(512)
    x = torch.randn(900, 1, 256, generator=g)
    gy = torch.randn(900, 1, 512, generator=g)
    xo = x.clone().requires_grad_(True)
    F.layer_norm(lin(xo).relu() + 0.0, (512,), ln.weight, ln.bias).backward(gy)
    want = [xo.grad, lin.weight.grad.clone(), lin.bias.grad.clone()]
    yo1 = lin(x).relu()
    yo2 = F.layer_norm(lin(x), (512,), ln.weight, ln.bias)

    ling, lng = torch.nn.Linear(256, 512).cuda(), torch.nn.LayerNorm(512).cuda()
    ling.load_state_dict(lin.state_dict()); lng.load_state_dict(ln.state_dict())
    xg = x.cuda().requires_grad_(True)
    h = glue.linear_relu(xg, ling)
    assert _rel(h.detach(), yo1.detach()) <= TOL
    fused.add_layernorm(h, lng).backward(gy.cuda())
    assert _rel(xg.grad, want[0]) <= 5e-5
    assert _rel(ling.weight.grad, want[1]) <= 5e-5 and _rel(ling.bias.grad, want[2]) <= 5e-5

    ling.zero_grad(set_to_none=True)
    xg = x.cuda().requires_grad_(True)
    y2 = fused.add_layernorm(glue.fast_linear(xg, ling, add_bias=False), lng, xbias=ling.bias)
    assert _rel(y2.detach(), yo2.detach()) <= TOL
    lin.zero_grad(); xo = x.clone().requires_grad_(True)
    F.layer_norm(lin(xo), (512,), ln.weight, ln.bias).backward(gy)
    y2.backward(gy.cuda())
    assert _rel(xg.grad, xo.grad) <= 5e-5
    assert _rel(ling.weight.grad, lin.weight.grad) <= 5e-5 and _rel(ling.bias.grad, lin.bias.grad) <= 5e-5


def test_cat_linear_is_three_linears():
    """One GEMM over concatenated weights == the three generator Linears, gradients included
    (immediate and deferred), with zero padding columns that receive no gradient."""
    from graph_detr4d_b200 import glue
    g = torch.Generator().manual_seed(21)
    lins = [torch.nn.Linear(256, n) for n in (128, 96, 6)]
    x = torch.randn(1, 900, 256, generator=g)
    gy = torch.randn(1, 900, 232, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = torch.cat([l(xo) for l in lins], -1)
    yo.backward(gy[..., :230])
    lg = [torch.nn.Linear(256, n).cuda() for n in (128, 96, 6)]
    for a_, b_ in zip(lg, lins):
        a_.load_state_dict(b_.state_dict())
    for deferred in (False, True):
        for l in lg:
            l.zero_grad(set_to_none=True)
        xg = x.cuda().requires_grad_(True)
        y = glue.cat_linear(xg, lg, 232)
        assert tuple(y.shape) == (1, 900, 232) and float(y[..., 230:].abs().max()) == 0.0
        assert _rel(y[..., :230].detach(), yo.detach()) <= TOL
        if deferred:
            with DeferredWgrad() as wq:
                y.backward(gy.cuda())
                wq.flush()
        else:
            y.backward(gy.cuda())
        assert _rel(xg.grad, xo.grad) <= 5e-5
        for a_, b_ in zip(lg, lins):
            assert _rel(a_.weight.grad, b_.weight.grad) <= 5e-5 and _rel(a_.bias.grad, b_.bias.grad) <= 5e-5


def test_position_encoder_uses_fused_stages_and_matches_torch():
    seq = modules._position_encoder(3, 256)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 900, 3, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = seq(xo)
    yo.sum().backward()
    want = [p.grad.clone() for p in seq.parameters()]
    seq_g = modules._position_encoder(3, 256).cuda()
    seq_g.load_state_dict(seq.state_dict())
    xg = x.cuda().requires_grad_(True)
    from graph_detr4d_b200 import ops
    n0 = ops.launch_count()
    yg = modules._run_position_encoder(seq_g, xg)
    assert ops.launch_count() - n0 == 2                                # two fused LN+ReLU launches
    yg.sum().backward()
    assert _rel(yg.detach(), yo.detach()) <= TOL
    assert _rel(xg.grad, xo.grad) <= 1e-4
    for p, w in zip(seq_g.parameters(), want):
        assert _rel(p.grad, w) <= 1e-4


def test_glue_refuses_cpu_tensors():
    with pytest.raises(RuntimeError):
        fused.inverse_sigmoid(torch.rand(4, 3))
    with pytest.raises(RuntimeError):
        fused.add_layernorm(torch.randn(4, 256), torch.nn.LayerNorm(256))


def test_multi_tensor_adamw_matches_torch_adamw():
    """One-launch AdamW == torch.optim.AdamW over several steps, odd sizes and grad-less params included."""
    from graph_detr4d_b200.optim import MultiTensorAdamW
    g = torch.Generator().manual_seed(31)
    shapes = [(256, 256), (512,), (3, 256), (10,), (4097,), (1,), (900, 3), (7, 5, 3)]
    base = [torch.randn(s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(t.clone().cuda()) for t in base]
    pb = [torch.nn.Parameter(t.clone().cuda()) for t in base]
    skip = 3                                                              # never gets a gradient
    oa = torch.optim.AdamW(pa, lr=2e-3, weight_decay=0.05, fused=True)
    ob = MultiTensorAdamW(pb, lr=2e-3, weight_decay=0.05)
    for it in range(6):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == skip:
                continue
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it - 3))
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for i, (a, b) in enumerate(zip(pa, pb)):
        assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max()) + 1e-7, i
    assert torch.equal(pb[skip].detach().cpu(), base[skip])
    assert float(ob.step_t) == 6.0


@pytest.mark.parametrize("use", ["both", "y_only", "y2_only"])
def test_add_layernorm_second_output_with_pos(use):
    """(y, y + pos) in one launch; the backward sums the two incoming gradients in the kernel and
    the deferred gamma/beta reduction sees the summed gradient."""
    g = torch.Generator().manual_seed(41)
    C, rows = 256, 900
    x, r, pos = (torch.randn(rows, 1, C, generator=g) for _ in range(3))
    ln = torch.nn.LayerNorm(C)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(C, generator=g)); ln.bias.copy_(torch.randn(C, generator=g))
    g1, g2 = torch.randn(rows, 1, C, generator=g), torch.randn(rows, 1, C, generator=g)

    def loss(y, y2):
        return {"both": (y * g1.to(y.device)).sum() + (y2 * g2.to(y.device)).sum(),
                "y_only": (y * g1.to(y.device)).sum(), "y2_only": (y2 * g2.to(y.device)).sum()}[use]

    xo, ro, po = (t.clone().requires_grad_(True) for t in (x, r, pos))
    yo = F.layer_norm(xo + ro, (C,), ln.weight, ln.bias)
    loss(yo, yo + po).backward()
    lng = torch.nn.LayerNorm(C).cuda()
    lng.load_state_dict(ln.state_dict())
    for deferred in (False, True):
        lng.zero_grad(set_to_none=True)
        xg, rg, pg = (t.cuda().requires_grad_(True) for t in (x, r, pos))
        y, y2 = fused.add_layernorm(xg, lng, rg, pos=pg)
        assert _rel(y.detach(), yo.detach()) <= TOL and _rel(y2.detach(), (yo + po).detach()) <= TOL
        if deferred:
            with DeferredWgrad() as wq:
                loss(y, y2).backward()
                wq.flush()
        else:
            loss(y, y2).backward()
        assert _rel(xg.grad, xo.grad) <= 5e-5 and _rel(rg.grad, ro.grad) <= 5e-5
        if use == "y_only":
            assert pg.grad is None
        else:
            assert _rel(pg.grad, po.grad) <= TOL
        assert _rel(lng.weight.grad, ln.weight.grad) <= 5e-5 and _rel(lng.bias.grad, ln.bias.grad) <= 5e-5


@pytest.mark.parametrize("rows,cols", [(8 * 900, 900), (37, 64), (5, 1024), (3, 4)])
def test_one_pass_softmax_backward_matches_aten(rows, cols):
    """gd4d_softmax_bwd (in place over the incoming gradient) vs torch._softmax_backward_data."""
    from graph_detr4d_b200 import fused
    g = torch.Generator().manual_seed(rows + cols)
    p = torch.randn(rows, cols, generator=g).cuda().softmax(-1)
    go = torch.randn(rows, cols, generator=g).cuda()
    want = torch._softmax_backward_data(go, p, -1, p.dtype)
    assert fused.can_fuse_softmax_bwd(go, p)
    got = fused.softmax_bwd_(go.clone(), p)
    assert _rel(got, want.cpu()) <= 2e-6
    assert not fused.can_fuse_softmax_bwd(go[:, :cols - 1], p[:, :cols - 1])     # ragged / strided: ATen path


@pytest.mark.parametrize("with_pos,copies,used", [(False, 1, (True, True)), (True, 2, (True, True, True, True)),
                                                  (True, 2, (False, True, False, True)), (False, 2, (False, False, True))])
def test_add_layernorm_output_copies_sum_their_gradients_in_the_kernel(with_pos, copies, used):
    """``copies``: every consumer of a fused LayerNorm result gets its own identical output; whichever subset of the
    outputs receives a gradient, dX / dgamma / dbeta equal those of ONE output consumed by all of them."""
    g = torch.Generator().manual_seed(11)
    x, r, pos = (torch.randn(50, 2, 256, generator=g) for _ in range(3))
    ln = torch.nn.LayerNorm(256)
    with torch.no_grad():
        ln.weight.add_(torch.randn(256, generator=g) * 0.1)
        ln.bias.add_(torch.randn(256, generator=g) * 0.1)
    n_out = 1 + int(with_pos) + copies
    gys = [torch.randn(50, 2, 256, generator=g) for _ in range(n_out)]
    # reference: plain torch, one y consumed by everything
    xr, rr, pr = (t.clone().requires_grad_(True) for t in (x, r, pos))
    y = torch.nn.functional.layer_norm(xr + rr, (256,), ln.weight, ln.bias, ln.eps)
    outs_ref = [y] + ([y + pr] if with_pos else []) + [y] * copies
    sum((o * gy).sum() for o, gy, u in zip(outs_ref, gys, used) if u).backward()
    want = (xr.grad, rr.grad, ln.weight.grad.clone(), ln.bias.grad.clone())
    ln.zero_grad()
    lng = ln.cuda()
    xg, rg, pg = (t.cuda().requires_grad_(True) for t in (x, r, pos))
    outs = fused.add_layernorm(xg, lng, rg, pos=pg if with_pos else None, copies=copies)
    assert len(outs) == n_out
    for o in outs[1 + int(with_pos):]:
        assert torch.equal(o, outs[0]) and o.data_ptr() != outs[0].data_ptr()
    sum((o * gy.cuda()).sum() for o, gy, u in zip(outs, gys, used) if u).backward()
    for a, b in zip((xg.grad, rg.grad, lng.weight.grad, lng.bias.grad), want):
        assert _rel(a, b) <= 2e-5
    if with_pos and used[1]:
        assert _rel(pg.grad, pr.grad) <= 1e-6
