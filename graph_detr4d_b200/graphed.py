"""CUDA-graph capture of the decoder training step.

At batch 1 / 900 queries the decoder step is ~1100 tiny launches and the host
(Python + autograd dispatch) cannot issue them as fast as a B200 retires them, so
the whole step -- forward, backward (our fused kernels included: they are plain
stream-ordered launches through the C ABI, hence capturable) and the optimizer --
is captured once and replayed.  No tracing compiler is involved: capture records
exactly the kernels eager mode launched.

Data-parallel use (one process per GPU): gradients live in ONE flat fp32 buffer
(every ``param.grad`` is a view of it), so the only collective of the step is a
single NCCL all-reduce of that buffer between the captured forward/backward and
the captured optimizer step (the sampling path itself needs no
communication).  (r1: capturing the NCCL all-reduce INSIDE the step graph was tried and
hung during capture with torch 2.11 / NCCL 2.28 on 2 GPUs; the collective therefore stays an
eager, stream-ordered call between the two graphs.)
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import decoder as _decoder
from . import modules
from .glue import DeferredWgrad
from .optim import MultiTensorAdamW

MULTI_TENSOR_ADAMW = os.environ.get("GD4D_MULTI_ADAMW", "1") != "0"      # A/B switches for measurements
COALESCED_ALLREDUCE = os.environ.get("GD4D_COALESCED_ALLREDUCE", "1") != "0"


class HostFeatureBuffer:
    """The feature maps of one step in ONE pinned host allocation (``views`` are the per-level
    (B,N,C,H,W) tensors inside it), so that a step's host->device transfer is a single
    ``cudaMemcpyAsync`` instead of one per level.

    ``dtype`` is the WIRE format.  It may be narrower than the dtype the step computes on: the
    reference's backbone + FPN run under fp16 autocast and hand the head ``.float()`` copies
    (detectors/detr3d.py:68, ``auto_fp16(apply_to=('img'), out_fp32=True)``; 24 of its 29 configs set
    ``fp16 = dict(loss_scale=512.)``), so the fp32 maps the decoder consumes are fp16-exact and a
    ``torch.float16`` host buffer carries them losslessly at half the PCIe bytes; ``commit`` widens them
    on the device.  (Measured on the 8-GPU box, profiles/r2_h2d_probe_n8.json: with 8 ranks copying at
    once GPUs 0-3 get 22.5 GB/s each, so 189 MB of fp32 maps cost 8.4 ms per step -- more than the
    4.9 ms step itself.)"""

    def __init__(self, shapes: Sequence[Sequence[int]], dtype: torch.dtype = torch.float32):
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        self.flat = torch.empty(sum(sizes), dtype=dtype).pin_memory()
        self.views, o = [], 0
        for s, n in zip(shapes, sizes):
            self.views.append(self.flat[o:o + n].view(*s))
            o += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


def _flat_like(feats: Sequence[torch.Tensor]):
    """One device allocation holding copies of ``feats`` back to back -> (flat, per-tensor views)."""
    flat = torch.empty(sum(f.numel() for f in feats), device=feats[0].device, dtype=feats[0].dtype)
    views, o = [], 0
    for f in feats:
        views.append(flat[o:o + f.numel()].view(f.shape))
        o += f.numel()
    return flat, views


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, forward_loss: Callable[[Sequence[torch.Tensor]], torch.Tensor],
                 example_feats: Sequence[torch.Tensor], img_metas, lr=2e-4, weight_decay=0.01,
                 warmup_iters: int = 3, feats_require_grad: bool = True, world_size: int = 1):
        """``forward_loss(feats) -> scalar loss`` must only launch capturable work."""
        self.model = model
        self.world = world_size
        dev = example_feats[0].device
        self.device = dev
        params = [p for p in model.parameters() if p.requires_grad]
        self.params = params
        n = sum(p.numel() for p in params)
        # Autograd is left to hand each parameter its freshly computed gradient (``p.grad = None``
        # before backward), which costs no kernel; accumulating into pre-set ``.grad`` views instead
        # costs one elementwise add PER PARAMETER per step (~250 launches, ~0.7 ms of this step in r1).
        # N > 1: almost every gradient is a view into one of ~15 large batched result buffers of
        # DeferredWgrad.flush(); those buffers (plus the few directly produced gradients) are
        # all-reduced IN PLACE by one grouped NCCL launch.  (GD4D_COALESCED_ALLREDUCE=0: the r1
        # scheme -- gather every gradient into ONE flat buffer with a multi-tensor copy, all-reduce it.)
        flat = world_size > 1 and not COALESCED_ALLREDUCE
        self.flat_grad = torch.zeros(n if flat else 1, device=dev, dtype=torch.float32)
        self.flat_views = []
        self._reduce: List[torch.Tensor] = []
        self._grouped_ok = True
        if flat:
            o = 0
            for p in params:
                self.flat_views.append(self.flat_grad[o:o + p.numel()].view_as(p))
                o += p.numel()
        # one-launch AdamW over all parameter tensors (optim.py); torch.optim.AdamW arithmetic
        self.opt = (MultiTensorAdamW(params, lr=lr, weight_decay=weight_decay) if MULTI_TENSOR_ADAMW else
                    torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, fused=True, capturable=True))
        # the static inputs of the captured step live back to back in ONE allocation, so a new step's
        # maps arrive with one copy (host->device or device->device) instead of one per level
        same = len({f.dtype for f in example_feats}) == 1
        if same:
            self._static_flat, views = _flat_like(example_feats)
            with torch.no_grad():
                for v, f in zip(views, example_feats):
                    v.copy_(f)
            self.static_feats: List[torch.Tensor] = [v.requires_grad_(feats_require_grad) for v in views]
        else:
            self._static_flat = None
            self.static_feats = [f.detach().clone().requires_grad_(feats_require_grad) for f in example_feats]
        self.img_metas = img_metas
        modules.lidar2img_device(img_metas, dev)           # upload once, outside capture
        self._forward_loss = forward_loss

        def fwd_bwd():
            _decoder.STATIC_GENERATOR_PACKS = True         # forward and backward alternate strictly here
            modules.clear_pack_cache()                     # the pack kernels must be part of the capture
            for p in params:
                p.grad = None
            for f in self.static_feats:
                f.grad = None
            loss = forward_loss(self.static_feats)
            with DeferredWgrad() as wq:                    # weight grads of the small Linears: batched GEMMs
                loss.backward()
                wq.flush()
            if self.world > 1:
                if COALESCED_ALLREDUCE:
                    # all-reduce the few large batched result buffers in place, plus the handful of
                    # gradients autograd produced directly: no gather copy, the optimizer reads p.grad
                    self._reduce = list(wq.buffers) + [p.grad for p in params
                                                       if p.grad is not None and id(p) not in wq.covered]
                else:
                    have = [(v, p.grad) for v, p in zip(self.flat_views, params) if p.grad is not None]
                    torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
                    for v, p in zip(self.flat_views, params):
                        p.grad = v                         # the optimizer reads the (all-reduced) flat views
            _decoder.STATIC_GENERATOR_PACKS = False
            return loss.detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup_iters):
                fwd_bwd()
                self._allreduce()
                self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb):
            self.static_loss = fwd_bwd()
        if MULTI_TENSOR_ADAMW:
            self.opt.prepare()                             # pointer table of the CAPTURED gradient tensors
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt):
            self.opt.step()
        torch.cuda.synchronize(dev)

    def _allreduce(self):
        if self.world <= 1:
            return
        if not COALESCED_ALLREDUCE:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
            return
        # ONE grouped NCCL launch (ncclGroupStart/End) over ~20 tensors, in place
        if self._grouped_ok:
            try:
                with dist.distributed_c10d._coalescing_manager():
                    for t in self._reduce:
                        dist.all_reduce(t, op=dist.ReduceOp.AVG)
                return
            except (RuntimeError, ValueError, AttributeError, AssertionError) as e:   # private torch API: degrade, loudly
                import warnings
                warnings.warn(f"grouped all-reduce unavailable ({e!r}); falling back to one all-reduce per buffer")
                self._grouped_ok = False
                dist.distributed_c10d._world.pg_coalesce_state.pop(dist.distributed_c10d._get_default_group(), None)
        for t in self._reduce:
            dist.all_reduce(t, op=dist.ReduceOp.AVG)

    def set_inputs(self, feats: Optional[Sequence[torch.Tensor]] = None, img_metas=None):
        """Stream-ordered refresh of the static inputs (H2D when ``feats`` are host tensors)."""
        if img_metas is not None:
            modules.lidar2img_device(img_metas, self.device)        # in-place update of the static buffer
        if feats is not None:
            with torch.no_grad():
                for dst, src in zip(self.static_feats, feats):
                    dst.copy_(src, non_blocking=True)

    # ---- host-buffer pipeline: H2D of step i+1 overlaps the compute of step i ------------
    def prefetch(self, feats_host):
        """Start the H2D copy of the NEXT step's (pinned) feature maps on a copy stream.
        ``feats_host``: a ``HostFeatureBuffer`` (ONE copy) or a sequence of pinned tensors (one per level)."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            # the staging buffer has the WIRE dtype of the host buffer (fp16 maps stay fp16 until commit)
            wire = feats_host.flat.dtype if isinstance(feats_host, HostFeatureBuffer) else None
            if self._static_flat is not None:
                like = [f.detach() if wire is None else torch.empty(f.shape, device=self.device, dtype=wire)
                        for f in self.static_feats]
                self._staging_flat, self._staging = _flat_like(like)
            else:
                self._staging_flat, self._staging = None, [torch.empty_like(f) for f in self.static_feats]
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(self.device))
        self._copy_stream.wait_event(self._consumed)        # staging is free again
        with torch.cuda.stream(self._copy_stream), torch.no_grad():
            if isinstance(feats_host, HostFeatureBuffer) and self._staging_flat is not None and \
                    feats_host.flat.dtype == self._staging_flat.dtype and \
                    feats_host.flat.numel() == self._staging_flat.numel():
                self._staging_flat.copy_(feats_host.flat, non_blocking=True)      # ONE cudaMemcpyAsync
            else:
                srcs = feats_host.views if isinstance(feats_host, HostFeatureBuffer) else feats_host
                for dst, src in zip(self._staging, srcs):
                    dst.copy_(src, non_blocking=True)
            self._staged.record(self._copy_stream)

    def reset_pipeline(self):
        """Forget the staging buffers (the next ``prefetch`` re-creates them for its host buffer's wire dtype)."""
        if hasattr(self, "_copy_stream"):
            torch.cuda.synchronize(self.device)
            for name in ("_copy_stream", "_staging_flat", "_staging", "_staged", "_consumed"):
                delattr(self, name)

    def commit(self, img_metas=None):
        """Make the prefetched maps the inputs of the next ``step()`` (one D2D copy, which also widens a
        narrower wire dtype to the compute dtype)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged)
        if self._staging_flat is not None:
            if img_metas is not None:
                modules.lidar2img_device(img_metas, self.device)
            with torch.no_grad():
                # one D2D for all levels; widens fp16 / bf16 wire maps to the compute dtype on the way
                self._static_flat.copy_(self._staging_flat, non_blocking=True)
        else:
            self.set_inputs(self._staging, img_metas)
        self._consumed.record(cur)

    def step(self) -> torch.Tensor:
        self.graph_fb.replay()
        self._allreduce()
        self.graph_opt.replay()
        return self.static_loss
