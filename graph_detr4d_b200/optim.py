"""AdamW for the decoder's ~200 small parameter tensors as ONE launch
(include/gd4d_glue.h: gd4d_adamw_multi; csrc/glue.cu).

torch's fused AdamW hands each CTA a 65536-element chunk, so ~7 M parameters spread over
~200 tensors become ~150 CTAs in 6 launches (39 us each on a B200, r1 profile); here a CTA
owns 4096 elements and the whole update is one HBM-bound launch.  Same arithmetic as
``torch.optim.AdamW(fused=True, capturable=True)``: decoupled weight decay, bias corrections
from a device-side step counter, so ``step()`` is CUDA-graph capturable.

The kernel reads a device table of (param, grad, exp_avg, exp_avg_sq) pointers.  It is rebuilt
(one small pinned H2D copy) whenever a gradient tensor's address changed since the previous
step -- every step in eager mode, never inside a captured graph, where autograd's allocations
are replayed at fixed addresses: call ``prepare()`` once after the backward has been captured.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _lib
from .ops import _count, _stream_ptr


class MultiTensorAdamW:
    def __init__(self, params: Sequence[torch.nn.Parameter], lr=2e-4, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.01):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise TypeError("MultiTensorAdamW needs contiguous float32 CUDA parameters on one device")
        self.device = dev
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        n = sum(p.numel() for p in self.params)
        # 4-element alignment of every tensor's slice keeps the 16-byte vector path
        offs, o = [], 0
        for p in self.params:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self._state = torch.zeros(2, o, device=dev, dtype=torch.float32)        # exp_avg, exp_avg_sq
        self.exp_avg = [self._state[0, a:a + p.numel()].view_as(p) for a, p in zip(offs, self.params)]
        self.exp_avg_sq = [self._state[1, a:a + p.numel()].view_as(p) for a, p in zip(offs, self.params)]
        self.step_t = torch.zeros((), device=dev, dtype=torch.float32)
        self._chunk = _lib.load().gd4d_adamw_chunk()
        self._grad_ptrs = None
        self._table_dev = self._block_map = None
        self.n_blocks = 0
        self.numel = n

    def prepare(self):
        """(Re)build the device pointer table and block map from the parameters' current ``.grad``
        tensors (parameters without a gradient are skipped, like torch.optim)."""
        ptrs = [None if p.grad is None else p.grad.data_ptr() for p in self.params]
        if ptrs == self._grad_ptrs:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("gradient addresses changed during CUDA-graph capture: call prepare() on the "
                               "captured backward's gradients before capturing step()")
        rows, bmap = [], []
        for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq):
            g = p.grad
            if g is None:
                continue
            if g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
                raise RuntimeError("gradients must be contiguous float32 tensors shaped like their parameters")
            bmap += [(len(rows), c) for c in range((p.numel() + self._chunk - 1) // self._chunk)]
            rows.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()))
        if not rows:
            raise RuntimeError("MultiTensorAdamW.step(): no parameter has a gradient")
        # fresh device tensors every rebuild: a launch still in flight keeps reading the old ones
        self._table_dev = torch.tensor(rows, dtype=torch.int64).to(self.device)
        self._block_map = torch.tensor(bmap, dtype=torch.int32).to(self.device)
        self.n_blocks = len(bmap)
        self._grad_ptrs = ptrs

    @torch.no_grad()
    def step(self):
        self.prepare()
        self.step_t += 1.0
        st = _lib.load().gd4d_adamw_multi(
            self._table_dev.data_ptr(), self._block_map.data_ptr(), self.n_blocks, self.step_t.data_ptr(),
            self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, _stream_ptr(self.device))
        _lib.check(st, "gd4d_adamw_multi")
        _count()

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
