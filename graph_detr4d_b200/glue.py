"""Library-GEMM glue around the hot path, tuned for the
decoder's shape (M = B*Q = 900 rows): stock torch picks a 2-CTA reduction for every
Linear's bias gradient (13 us each, ~70 per step) and a memory-efficient attention
kernel built for long sequences (310 us fwd+bwd for 900 queries).  Same maths,
cuBLAS underneath; nothing here touches the sampling path.
"""
from __future__ import annotations

import math

import torch


class DeferredWgrad:
    """Delayed, BATCHED weight gradients for the decoder's many small Linears.

    Every Linear of the decoder sees M = B*Q = 900 rows, so each weight gradient is a
    (N x 900) @ (900 x K) GEMM that cuBLAS runs on ~64 CTAs for 16 us (80 of them per step:
    1.25 ms in r1).  Inside ``with DeferredWgrad():`` the backward of ``fast_linear`` only
    computes dX and queues (grad_out, input); ``flush()`` -- called once after
    ``loss.backward()`` -- groups the queue by shape and computes each group's weight
    gradients with ONE batched GEMM (thousands of CTAs) and its bias gradients with one
    reduction, then assigns ``param.grad``.  Same maths; only for leaf Parameters, only
    when autograd hooks on parameter gradients are not needed (no DDP): GraphedTrainStep."""

    _active = None

    def __init__(self):
        self.items = []        # (weight_param, bias_param, row0, row1, g2, x2)
        self.ln_items = []     # (weight_param, bias_param, g2, x2, mean, rstd)
        # after flush(): the few large contiguous result buffers, and the ids of the parameters whose
        # .grad is a view into one of them -- a data-parallel caller all-reduces THESE instead of
        # gathering ~200 gradient tensors into a flat buffer first (graphed.GraphedTrainStep)
        self.buffers = []
        self.covered = set()

    def __enter__(self):
        DeferredWgrad._active = self
        return self

    def __exit__(self, *exc):
        DeferredWgrad._active = None
        return False

    def flush(self):
        groups = {}
        for it in self.items:
            g2, x2 = it[4], it[5]
            groups.setdefault((g2.shape[0], g2.shape[1], x2.shape[1]), []).append(it)
        partial = {}
        gid = 0
        for (_, _, _), its in groups.items():
            gid += 1
            # row-slice items (packed parameters) first, in encounter order: their results are then
            # consecutive slices of dW / dB in the same module order in every group (see below)
            its.sort(key=lambda it: 0 if not (it[2] == 0 and it[3] == it[0].shape[0]) else 1)
            G = torch.stack([it[4] for it in its])                     # (n, M, N)
            X = torch.stack([it[5] for it in its])                     # (n, M, K)
            dW = torch.bmm(G.transpose(1, 2), X)                       # (n, N, K)
            ones = _ones_row(G, G.shape[1]).expand(G.shape[0], 1, G.shape[1])
            dB = torch.bmm(ones, G).squeeze(1)                         # (n, N): GEMV, not aten::sum (2-CTA reduce)
            n_part = sum(1 for it in its if not (it[2] == 0 and it[3] == it[0].shape[0]))
            if n_part < len(its):
                self.buffers += [dW[n_part:], dB[n_part:]]             # the whole-parameter results (contiguous)
            for i, (wp, bp, r0, r1, _, _) in enumerate(its):
                whole = r0 == 0 and r1 == wp.shape[0]
                if whole:
                    self._assign(wp, dW[i])
                    if bp is not None:
                        self._assign(bp, dB[i])
                else:                                                  # row slice of a packed parameter
                    partial.setdefault(id(wp), (wp, bp, []))[2].append((r0, r1, dW[i], dB[i], gid, i, dW, dB))
        # Packed parameters whose row slices tile them exactly (nn.MultiheadAttention.in_proj:
        # [q,k rows | v rows]) and sit at CONSECUTIVE indices of the same batched results across
        # modules (layer after layer): ONE cat along the row axis for all of them.
        batched = {}
        for key, (wp, bp, parts) in partial.items():
            parts.sort(key=lambda t: t[0])
            tiles = parts[0][0] == 0 and parts[-1][1] == wp.shape[0] and \
                all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sig = (tiles, bp is not None, tuple((t[0], t[1], t[4]) for t in parts))
            batched.setdefault(sig, []).append(key)
        done = set()
        for sig, keys in batched.items():
            if not sig[0] or len(keys) < 2:
                continue
            plists = [partial[k][2] for k in keys]
            nparts = len(plists[0])
            idx = [[pl[j][5] for pl in plists] for j in range(nparts)]          # per part: batch indices
            if not all(ix == list(range(ix[0], ix[0] + len(ix))) or ix == list(range(ix[0], ix[0] - len(ix), -1))
                       for ix in idx):
                continue
            def take(t, ix):
                lo, hi = min(ix), max(ix) + 1
                v = t[lo:hi]
                return v if ix[0] == lo else v.flip(0)
            gw_all = torch.cat([take(plists[0][j][6], idx[j]) for j in range(nparts)], dim=1)   # (m, rows, K)
            gb_all = torch.cat([take(plists[0][j][7], idx[j]) for j in range(nparts)], dim=1) if sig[1] else None
            self.buffers += [gw_all] + ([gb_all] if gb_all is not None else [])
            for m, k in enumerate(keys):
                wp, bp, _ = partial[k]
                self._assign(wp, gw_all[m])
                if bp is not None:
                    self._assign(bp, gb_all[m])
                done.add(k)
        for key, (wp, bp, parts) in partial.items():
            if key in done:
                continue
            parts = [t[:4] for t in parts]
            tiles = parts[0][0] == 0 and parts[-1][1] == wp.shape[0] and \
                all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            if tiles:                                                  # [qk rows | v rows] -> one cat each
                gw = torch.cat([t[2] for t in parts])
                gb = torch.cat([t[3] for t in parts]) if bp is not None else None
            else:
                gw = torch.zeros_like(wp)
                gb = torch.zeros_like(bp) if bp is not None else None
                for r0, r1, dw, db in parts:
                    gw[r0:r1] += dw
                    if gb is not None:
                        gb[r0:r1] += db
            wp.grad = gw if wp.grad is None else wp.grad + gw
            if bp is not None:
                bp.grad = gb if bp.grad is None else bp.grad + gb
        self.items.clear()
        # LayerNorm gamma/beta: d_gamma = sum_rows g * xhat, d_beta = sum_rows g, batched by width
        ln_groups = {}
        for it in self.ln_items:
            ln_groups.setdefault((it[2].shape[0], it[2].shape[1]), []).append(it)
        for its in ln_groups.values():
            G = torch.stack([it[2] for it in its])                     # (n, M, C)
            X = torch.stack([it[3] for it in its])
            mean = torch.stack([it[4] for it in its])                  # (n, M, 1)
            rstd = torch.stack([it[5] for it in its])
            ones = _ones_row(G, G.shape[1]).expand(G.shape[0], 1, G.shape[1])
            dG = torch.bmm(ones, G * ((X - mean) * rstd)).squeeze(1)   # (n, C)
            dB = torch.bmm(ones, G).squeeze(1)
            self.buffers += [dG, dB]
            for i, (wp, bp, *_rest) in enumerate(its):
                self._assign(wp, dG[i])
                self._assign(bp, dB[i])
        self.ln_items.clear()

    def _assign(self, param, grad_view):
        """param.grad = view of a recorded buffer (covered), or accumulate (then no longer covered)."""
        if param.grad is None:
            param.grad = grad_view
            self.covered.add(id(param))
        else:
            param.grad = param.grad + grad_view
            self.covered.discard(id(param))


def _ones_row(like: torch.Tensor, n: int) -> torch.Tensor:
    """(1,1,n) row of ones: the GEMV that replaces aten::sum for bias gradients."""
    return _const(("ones", like.device, like.dtype, n),
                  lambda: torch.ones(1, 1, n, device=like.device, dtype=like.dtype))


class _FastLinearFn(torch.autograd.Function):
    """F.linear whose bias gradient is a (1,M)x(M,N) GEMM instead of aten::sum, and whose
    weight/bias gradients can be deferred to a batched GEMM (DeferredWgrad).
    ``owner`` = (weight_param, bias_param, row0, row1) when ``weight``/``bias`` are (row
    slices of) leaf Parameters."""

    @staticmethod
    def forward(ctx, x, weight, bias, owner=None, add_bias=True):
        """``add_bias=False``: the GEMM runs without its bias epilogue (a separate kernel in fp32
        cuBLAS) because the consumer adds the bias itself (fused.add_layernorm ``xbias``); the
        bias gradient is still produced here -- it is the column sum of the same ``g``."""
        x2 = x.reshape(-1, x.shape[-1])
        ctx.save_for_backward(x2, weight)
        ctx.xshape = x.shape
        ctx.owner = owner
        y = x.new_empty(*x.shape[:-1], weight.shape[0])       # final shape: not a view, so a
        if add_bias:                                          # following in-place ReLU is legal
            torch.addmm(bias, x2, weight.t(), out=y.view(-1, weight.shape[0]))
        else:
            torch.mm(x2, weight.t(), out=y.view(-1, weight.shape[0]))
        return y

    @staticmethod
    def backward(ctx, g):
        x2, weight = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (g2 @ weight).view(ctx.xshape)
        q = DeferredWgrad._active
        if q is not None and ctx.owner is not None and ctx.needs_input_grad[1]:
            q.items.append((*ctx.owner, g2, x2))
            return dx, None, None, None, None
        if ctx.needs_input_grad[1]:
            dw = g2.t() @ x2
        if ctx.needs_input_grad[2]:
            db = (_ones_row(g2, g2.shape[0])[0] @ g2).view(-1)
        return dx, dw, db, None, None


class _LinearReluFn(torch.autograd.Function):
    """relu(F.linear(x)) as GEMM + ONE in-place bias+ReLU launch (gd4d_bias_act) instead of
    GEMM + cuBLAS bias epilogue kernel + ReLU kernel; backward masks the gradient once and then
    behaves like _FastLinearFn (deferred batched weight/bias gradients)."""

    @staticmethod
    def forward(ctx, x, weight, bias, owner=None):
        from . import fused
        x2 = x.reshape(-1, x.shape[-1])
        y = x.new_empty(*x.shape[:-1], weight.shape[0])
        torch.mm(x2, weight.t(), out=y.view(-1, weight.shape[0]))
        fused.bias_act_(y, bias, relu=True)
        ctx.save_for_backward(x2, weight, y)
        ctx.xshape = x.shape
        ctx.owner = owner
        return y

    @staticmethod
    def backward(ctx, g):
        x2, weight, y = ctx.saved_tensors
        g2 = torch.ops.aten.threshold_backward(g, y, 0).reshape(-1, g.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (g2 @ weight).view(ctx.xshape)
        q = DeferredWgrad._active
        if q is not None and ctx.owner is not None and ctx.needs_input_grad[1]:
            q.items.append((*ctx.owner, g2, x2))
            return dx, None, None, None
        if ctx.needs_input_grad[1]:
            dw = g2.t() @ x2
        if ctx.needs_input_grad[2]:
            db = (_ones_row(g2, g2.shape[0])[0] @ g2).view(-1)
        return dx, dw, db, None


def _pad_zeros(like: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    return _const(("pad", like.device, like.dtype, rows, cols),
                  lambda: torch.zeros(rows, cols, device=like.device, dtype=like.dtype))


class _CatLinearFn(torch.autograd.Function):
    """Several Linears on the SAME input as ONE GEMM over their concatenated weights:
    y (.., width) = x @ cat(W_i)^T + cat(b_i), columns padded with zeros to ``width``.
    Backward: one dX GEMM; each Linear's weight/bias gradient is its column block of G^T X
    (deferred and batched under DeferredWgrad).  inputs: x, width, packed, *(w0, b0, w1, b1, ...);
    ``packed`` = (W_cat, b_cat) if the caller keeps an up-to-date packed copy of the weights
    (decoder.py refreshes all layers' copies with one multi-tensor launch per forward), else None."""

    @staticmethod
    def forward(ctx, x, width: int, packed, *wb):
        from . import fused
        ws, bs = wb[0::2], wb[1::2]
        n = sum(int(w.shape[0]) for w in ws)
        K = x.shape[-1]
        pad = width - n
        x2 = x.reshape(-1, K)
        if packed is not None:
            wc, bc = packed
        else:
            wc = torch.cat(list(ws) + ([_pad_zeros(x, pad, K)] if pad else []))
            bc = torch.cat(list(bs) + ([_pad_zeros(x, 1, pad).view(-1)] if pad else []))
        y = x.new_empty(*x.shape[:-1], width)
        torch.mm(x2, wc.t(), out=y.view(-1, width))
        fused.bias_act_(y, bc, relu=False)
        ctx.save_for_backward(x2, wc)
        ctx.xshape = x.shape
        ctx.params = [(w, b) for w, b in zip(ws, bs)]
        return y

    @staticmethod
    def backward(ctx, g):
        x2, wc = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        dx = (g2 @ wc).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        q = DeferredWgrad._active
        grads, c0 = [], 0
        dwc = dbc = None
        for i, (w, b) in enumerate(ctx.params):
            c1 = c0 + int(w.shape[0])
            need = ctx.needs_input_grad[3 + 2 * i]
            if need and q is not None and w.is_leaf:
                q.items.append((w, b, 0, int(w.shape[0]), g2[:, c0:c1], x2))
                grads += [None, None]
            elif need:
                if dwc is None:
                    dwc = g2.t() @ x2
                    dbc = (_ones_row(g2, g2.shape[0])[0] @ g2).view(-1)
                grads += [dwc[c0:c1], dbc[c0:c1]]
            else:
                grads += [None, None]
            c0 = c1
        return (dx, None, None, *grads)


def cat_linear(x: torch.Tensor, lins, width: int, packed=None) -> torch.Tensor:
    """(.., width) = [lin_0(x) | lin_1(x) | ... | 0-padding], one GEMM (CUDA fp32 only)."""
    wb = []
    for lin in lins:
        wb += [lin.weight, lin.bias]
    return _CatLinearFn.apply(x, int(width), packed, *wb)


class _FastLayerNormFn(torch.autograd.Function):
    """LayerNorm whose gamma/beta gradients (a 12 us column reduction each, 30 per step) can be
    deferred to one batched reduction (DeferredWgrad); dX is computed immediately."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, owner):
        out, mean, rstd = torch.native_layer_norm(x, (x.shape[-1],), weight, bias, eps)
        ctx.save_for_backward(x, mean, rstd, weight, bias)
        ctx.owner = owner
        return out

    @staticmethod
    def backward(ctx, g):
        x, mean, rstd, weight, bias = ctx.saved_tensors
        q = DeferredWgrad._active
        g = g.contiguous()
        if q is not None and ctx.owner is not None:
            dx = torch.ops.aten.native_layer_norm_backward(g, x, (x.shape[-1],), mean, rstd, weight, bias,
                                                           [True, False, False])[0]
            C = x.shape[-1]
            q.ln_items.append((*ctx.owner, g.reshape(-1, C), x.reshape(-1, C), mean.reshape(-1, 1),
                               rstd.reshape(-1, 1)))
            return dx, None, None, None, None
        dx, dw, db = torch.ops.aten.native_layer_norm_backward(g, x, (x.shape[-1],), mean, rstd, weight, bias,
                                                               [True, True, True])
        return dx, dw, db, None, None


def fast_layer_norm(x: torch.Tensor, ln: torch.nn.LayerNorm) -> torch.Tensor:
    if not x.is_cuda or ln.weight is None or ln.bias is None or len(ln.normalized_shape) != 1:
        return ln(x)
    owner = (ln.weight, ln.bias) if ln.weight.requires_grad else None
    return _FastLayerNormFn.apply(x, ln.weight, ln.bias, ln.eps, owner)


def fast_linear(x: torch.Tensor, lin: torch.nn.Linear, add_bias: bool = True) -> torch.Tensor:
    if lin.bias is None or not x.is_cuda or x.dtype != lin.weight.dtype:
        assert add_bias, "deferred bias needs the CUDA path"
        return torch.nn.functional.linear(x, lin.weight, lin.bias)
    owner = (lin.weight, lin.bias, 0, lin.weight.shape[0]) if lin.weight.requires_grad else None
    return _FastLinearFn.apply(x, lin.weight, lin.bias, owner, add_bias)


def can_defer_bias(x: torch.Tensor, lin: torch.nn.Linear) -> bool:
    return lin.bias is not None and x.is_cuda and x.dtype == lin.weight.dtype == torch.float32


def linear_relu(x: torch.Tensor, lin: torch.nn.Linear) -> torch.Tensor:
    """relu(lin(x)); on CUDA fp32 the bias and the ReLU are one in-place launch after the GEMM."""
    from . import fused
    if not (fused.ENABLED and can_defer_bias(x, lin) and lin.out_features % 4 == 0):
        return torch.relu_(fast_linear(x, lin))
    owner = (lin.weight, lin.bias, 0, lin.weight.shape[0]) if lin.weight.requires_grad else None
    return _LinearReluFn.apply(x, lin.weight, lin.bias, owner)


_CONSTS = {}


def _const(key, make):
    """Persistent small constant tensors (no fill launch per use).  A constant first needed WHILE a
    CUDA graph is being captured is not cached: its memory belongs to that graph's private pool."""
    t = _CONSTS.get(key)
    if t is None:
        with torch.no_grad():
            t = make()
        if not (t.is_cuda and torch.cuda.is_current_stream_capturing()):
            _CONSTS[key] = t
    return t


def _zero(like: torch.Tensor) -> torch.Tensor:
    """(1,1,1) zero: baddbmm's ignored ``input`` (beta = 0)."""
    return _const(("zero", like.device, like.dtype),
                  lambda: torch.zeros(1, 1, 1, device=like.device, dtype=like.dtype))


class _ScaledBmmNT(torch.autograd.Function):
    """alpha * q @ k^T with alpha folded into the GEMMs (forward and both backward GEMMs): no
    separate elementwise scale of q (forward) or of its gradient (backward)."""

    @staticmethod
    def forward(ctx, q, k, alpha: float):
        ctx.save_for_backward(q, k)
        ctx.alpha = alpha
        return torch.baddbmm(_zero(q), q, k.transpose(1, 2), beta=0.0, alpha=alpha)

    @staticmethod
    def backward(ctx, g):
        q, k = ctx.saved_tensors
        z = _zero(g)
        dq = torch.baddbmm(z, g, k, beta=0.0, alpha=ctx.alpha) if ctx.needs_input_grad[0] else None
        dk = torch.baddbmm(z, g.transpose(1, 2), q, beta=0.0, alpha=ctx.alpha) if ctx.needs_input_grad[1] else None
        return dq, dk, None


class _AttnCoreFn(torch.autograd.Function):
    """softmax(alpha * q k^T) v for the packed projections qk (L,B,2E) and v (L,B,E), heads split
    as nn.MultiheadAttention does.  Every GEMM reads / writes its operands IN PLACE as strided
    batches (the per-head views of the packed buffers), forward and backward: no slice, transpose
    or concatenation copies, no zero fills (stock autograd: ~9 extra launches per layer).
    Batch 1 only (one scene per GPU): with B > 1 the (B,H) axes of the packed buffer do not merge
    into one batch stride and the caller uses the generic path."""

    @staticmethod
    def forward(ctx, qk, v, H: int, alpha: float):
        L, B, E2 = qk.shape
        E = E2 // 2
        d = E // H
        q = qk[..., :E].reshape(L, B * H, d).transpose(0, 1)              # (B*H,L,d) views
        k = qk[..., E:].reshape(L, B * H, d).transpose(0, 1)
        vv = v.reshape(L, B * H, d).transpose(0, 1)
        attn = torch.baddbmm(_zero(qk), q, k.transpose(1, 2), beta=0.0, alpha=alpha)
        attn = torch._softmax(attn, -1, False)
        out = qk.new_empty(L, B, E)
        torch.bmm(attn, vv, out=out.view(L, B * H, d).transpose(0, 1))
        ctx.save_for_backward(qk, v, attn)
        ctx.H, ctx.alpha = H, alpha
        return out

    @staticmethod
    def backward(ctx, g):
        qk, v, attn = ctx.saved_tensors
        L, B, E2 = qk.shape
        E = E2 // 2
        H, d = ctx.H, E // ctx.H
        g = g.contiguous()
        gg = g.view(L, B * H, d).transpose(0, 1)                            # (B*H,L,d)
        q = qk[..., :E].reshape(L, B * H, d).transpose(0, 1)
        k = qk[..., E:].reshape(L, B * H, d).transpose(0, 1)
        vv = v.reshape(L, B * H, d).transpose(0, 1)
        dv = torch.empty_like(v)
        torch.bmm(attn.transpose(1, 2), gg, out=dv.view(L, B * H, d).transpose(0, 1))
        from . import fused
        ds = torch.bmm(gg, vv.transpose(1, 2))
        if fused.can_fuse_softmax_bwd(ds, attn):
            fused.softmax_bwd_(ds, attn)                                  # one pass, in place
        else:
            ds = torch._softmax_backward_data(ds, attn, -1, attn.dtype)
        dqk = torch.empty_like(qk)
        z = _zero(qk)
        torch.baddbmm(z, ds, k, beta=0.0, alpha=ctx.alpha,
                      out=dqk[..., :E].view(L, B * H, d).transpose(0, 1))
        torch.baddbmm(z, ds.transpose(1, 2), q, beta=0.0, alpha=ctx.alpha,
                      out=dqk[..., E:].view(L, B * H, d).transpose(0, 1))
        return dqk, dv, None, None


def self_attention(query, query_pos, mha: torch.nn.MultiheadAttention, out_bias: bool = True, qk_in=None):
    """nn.MultiheadAttention(q=k=query+pos, v=query), batch_first=False, eval/dropout-free
    math: explicit bmm + softmax (faster than the flash/mem-efficient kernels at L=900,
    head_dim 32, fp32).  Uses the module's own packed parameters."""
    L, B, E = query.shape
    H = mha.num_heads
    d = E // H
    w, b = mha.in_proj_weight, mha.in_proj_bias
    if qk_in is None:                                   # else: the caller already has query + query_pos
        qk_in = query if query_pos is None else query + query_pos
    own = w.requires_grad
    qk = _FastLinearFn.apply(qk_in, w[:2 * E], b[:2 * E], (w, b, 0, 2 * E) if own else None, True)  # (L,B,2E)
    v = _FastLinearFn.apply(query, w[2 * E:], b[2 * E:], (w, b, 2 * E, 3 * E) if own else None, True)  # (L,B,E)
    if B == 1 and not (mha.dropout > 0 and mha.training) and qk.dtype == torch.float32:
        out = _AttnCoreFn.apply(qk, v, H, 1.0 / math.sqrt(d))             # (L,B,E)
        return fast_linear(out, mha.out_proj, add_bias=out_bias)
    q, k = qk[..., :E], qk[..., E:]
    q = q.reshape(L, B * H, d).transpose(0, 1)                          # (B*H,L,d)
    k = k.reshape(L, B * H, d).transpose(0, 1)
    v = v.reshape(L, B * H, d).transpose(0, 1)
    attn = _ScaledBmmNT.apply(q, k, 1.0 / math.sqrt(d)).softmax(-1)
    attn = torch.nn.functional.dropout(attn, mha.dropout, training=mha.training)   # eval: identity
    out = torch.bmm(attn, v).transpose(0, 1).reshape(L, B, E)
    return fast_linear(out, mha.out_proj, add_bias=out_bias)
