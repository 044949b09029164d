"""Host side of the cross-view sampling ops: torch tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory, the current CUDA stream, autograd
bookkeeping); all arithmetic of the path runs in libgd4d_xview.so.  There is no
CPU / eager fallback: CPU tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F32, MODE_A, MODE_C, MODE_V2, XViewParams

__all__ = ["XViewConfig", "GenLayout", "pack_features", "PackedFeatures", "xview_forward", "xview_backward",
           "xview_attention", "xview_attention_gen", "xview_forward_gen", "xview_backward_gen", "lidar2img_to_tensor", "prepare_forward", "prepare_backward",
           "launch_count", "MODE_A", "MODE_C", "MODE_V2"]

DYNAMIC_SCHEDULE = True   # persistent grid + work counter (False: one warp per item, static)
TMA_FORWARD = False       # wide forward through the cp.async.bulk staging path (xview_fwd_tma.cu)
L2_PREFETCH = os.environ.get("GD4D_L2_PREFETCH", "0") != "0"   # prefetch.global.L2 of the next batch's rows
# Wide backward: two implementations behind gd4d_xview_backward.
#   atomics (xview_bwd.cu, 1 launch): one 1 KB vector reduction per corner read into the shared grad map
#   sorted  (xview_bwd_sorted.cu, 5 launches): corner contributions sorted by pixel row, one warp owns a run:
#           value row read once, one reduction per run (486 k -> ~110 k reductions at N = 6)
# Measured r2 (profiles/r2_bwd_variants.json, bench layer-0 inputs, us per backward call, sorted vs atomics):
# fp32 N = 6: 136 vs 141, fp32 N = 12: 232 vs 271, bf16 N = 12: 254 vs 245; inside the graphed training step
# (tools/ab_step.py, same process, with the forward emitting the records): fp32 4.378 vs 4.448 ms at N = 6, 5.284 vs
# 5.585 ms at N = 12; bf16 4.312 vs 4.304 ms at N = 6, 5.161 vs 5.277 ms at N = 12.
# "auto" therefore takes the sorted path for fp32 maps, and for bf16 maps from 12 camera images up.
# GD4D_SORTED_BWD=0 / 1 forces one of them.
_SORTED_ENV = os.environ.get("GD4D_SORTED_BWD", "auto")
SORTED_BACKWARD = {"0": False, "1": True}.get(_SORTED_ENV, "auto")
_LAUNCHES = 0      # kernels of libgd4d_xview.so launched by this process (bench: gpu_launches)


def launch_count() -> int:
    return _LAUNCHES


def _count(n: int = 1):
    global _LAUNCHES
    _LAUNCHES += n


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"feature maps must be float32 or bfloat16, got {t.dtype}")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_SCHED = {}


def _sched_ptr(device) -> int:
    """Per-(device, stream) 2-word work counter for the persistent-grid scheduler.  Zeroed
    once; the kernels reset it themselves, so launches that share it must be ordered on one
    stream -- which is why it is keyed by the current stream."""
    key = (torch.device(device).index, _stream_ptr(device))
    t = _SCHED.get(key)
    if t is None:
        with torch.no_grad():
            t = torch.zeros(2, dtype=torch.int32, device=device)
        _SCHED[key] = t
    return t.data_ptr()


_BWD_WS = {}


def sorted_backward_active(mode: int, wide: bool, value_dtype_code: int, cams: int = 0) -> bool:
    """Which backward gd4d_xview_backward will run for this configuration (see SORTED_BACKWARD)."""
    if not (wide and mode == MODE_C):
        return False
    if SORTED_BACKWARD == "auto":
        return value_dtype_code == F32 or cams >= 12
    return bool(SORTED_BACKWARD)


def _attach_bwd_ws(p: XViewParams, device) -> int:
    """Scratch of the sorted wide backward, one per (device, stream) like the work counter: its
    counters + row histogram are zeroed once here, every launch leaves them zeroed again.
    Returns the number of kernels the backward call will launch."""
    if not sorted_backward_active(p.mode, bool(p.wide), p.value_dtype, p.B * p.N):
        return 1
    need = int(_lib.load().gd4d_xview_bwd_ws_bytes(C.byref(p)))
    if need < 0:
        _lib.check(need, "gd4d_xview_bwd_ws_bytes")
    key = (torch.device(device).index, _stream_ptr(device))
    t = _BWD_WS.get(key)
    rows = sum(p.B * p.N * p.level_h[l] * p.level_w[l] for l in range(p.L))
    head = 256 + (4 * rows + 255) // 256 * 256
    if t is None or t.numel() < need or t._gd4d_rows != rows:
        with torch.no_grad():
            t = torch.empty(need, dtype=torch.uint8, device=device)
            t[:head].zero_()
        t._gd4d_rows = rows          # a different row count moves the regions: start from a clean histogram
        _BWD_WS[key] = t
    p.bwd_ws = t.data_ptr()
    p.bwd_ws_bytes = t.numel()
    return 5


# Sorted backward, opt-in (GD4D_PRESORT=1): run its sort stage (emit, scan, scatter -- forward inputs only) right after
# the FORWARD kernel on a side stream, so the backward call is only the owner and finish kernels.  One scratch per
# (forward, backward) pair, owned by the autograd node; under CUDA-graph capture the side stream forks from and joins
# back into the capture through the events below (needs the node's backward inside the same capture).
# Measured r2 (tools/ab_step.py, same process): 4.599 vs 4.579 ms at N = 6, 5.292 vs 5.303 ms at N = 12 -- no gain:
# the 900-CTA emit kernel does not hide under the layer's GEMMs, it competes with them for the same SMs.  Off.
PRESORT = os.environ.get("GD4D_PRESORT", "0") != "0"
_SCRATCH_FREE = {}
# Sorted backward: let the FORWARD kernel emit the contribution records (GD4D_FLAG_FWD_EMIT: it builds the same
# per-item records anyway), so the backward call starts at the scan.  Scratch per (forward, backward) pair, owned by
# the autograd node, same stream throughout.  GD4D_FWD_EMIT=0 keeps the emit kernel in the backward.
FWD_EMIT = os.environ.get("GD4D_FWD_EMIT", "1") != "0"
_SORT_STREAMS = {}


def _presort(p: XViewParams, device):
    """``p``: parameters filled as for the backward (forward inputs are enough).  -> (scratch, done-event) or None."""
    if not (PRESORT and sorted_backward_active(p.mode, bool(p.wide), p.value_dtype, p.B * p.N)):
        return None
    lib = _lib.load()
    need = int(lib.gd4d_xview_bwd_ws_bytes(C.byref(p)))
    if need < 0:
        return None
    rows = sum(p.B * p.N * p.level_h[l] * p.level_w[l] for l in range(p.L))
    with torch.no_grad():
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        ws[:256 + (4 * rows + 255) // 256 * 256].zero_()
    p.bwd_ws, p.bwd_ws_bytes = ws.data_ptr(), ws.numel()
    dev = torch.device(device)
    side = _SORT_STREAMS.get(dev.index)
    if side is None:
        side = _SORT_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    side.wait_stream(main)                                  # inputs written, scratch head zeroed
    st = lib.gd4d_xview_backward_sort(C.byref(p), side.cuda_stream)
    _lib.check(st, "gd4d_xview_backward_sort")
    _count(3)
    done = torch.cuda.Event()
    done.record(side)
    if not torch.cuda.is_current_stream_capturing():
        ws.record_stream(side)
    return ws, done


def _emit_scratch(p: XViewParams, device, needs_grad: bool):
    """Scratch for a forward that emits the sorted backward's records: attaches it to ``p`` (forward params) and
    returns the ``presort`` tuple the backward takes, or None when that path does not apply."""
    if not (FWD_EMIT and not PRESORT and needs_grad and sorted_backward_active(p.mode, bool(p.wide), p.value_dtype, p.B * p.N)):
        return None
    need = int(_lib.load().gd4d_xview_bwd_ws_bytes(C.byref(p)))
    if need < 0:
        return None
    rows = sum(p.B * p.N * p.level_h[l] * p.level_w[l] for l in range(p.L))
    # a backward leaves its scratch clean (histogram and counters zero again), so scratches go back to a free list
    # when their backward has been launched and are handed out again without a memset; one whose backward never
    # ran is simply dropped.  Under CUDA-graph capture the warm-up steps have filled the list: no allocation and no
    # memset node inside the graph.
    key = (torch.device(device).index, need, rows)
    free = _SCRATCH_FREE.setdefault(key, [])
    if free:
        ws = free.pop()
        ev, stream_ptr = ws._gd4d_last
        if stream_ptr != _stream_ptr(device) and not torch.cuda.is_current_stream_capturing():
            if ev is not None:
                torch.cuda.current_stream(device).wait_event(ev)   # last used on another stream
    else:
        with torch.no_grad():
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            ws[:256 + (4 * rows + 255) // 256 * 256].zero_()
        ws._gd4d_key = key
    p.bwd_ws, p.bwd_ws_bytes = ws.data_ptr(), ws.numel()
    p.flags |= _lib.FLAG_FWD_EMIT
    return ws, None


def _recycle_scratch(presort):
    """After the backward has been launched (stream order keeps later users behind it): a forward-emit scratch is
    clean again and reusable.  Side-stream (presort) scratches are not pooled (their first use is on another stream)."""
    if presort is not None and presort[1] is None:
        key = getattr(presort[0], "_gd4d_key", None)
        if key is not None and len(_SCRATCH_FREE.setdefault(key, [])) < 64:
            ws = presort[0]
            if torch.cuda.is_current_stream_capturing():           # static inside a graph: same order every replay
                ws._gd4d_last = (None, _stream_ptr(ws.device))
            else:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(ws.device))
                ws._gd4d_last = (ev, _stream_ptr(ws.device))
            _SCRATCH_FREE[key].append(ws)


def clear_scratch_pool():
    """Drop the pooled forward-emit scratches (188 MB each at N = 6)."""
    _SCRATCH_FREE.clear()


def _use_presorted(p: XViewParams, presort, device) -> int:
    """Point ``p`` at a scratch prepared earlier: sorted by ``_presort`` on the side stream (wait for it), or holding
    the records the forward kernel emitted (same stream).  Returns the number of launches of the backward call."""
    ws, done = presort
    p.bwd_ws, p.bwd_ws_bytes = ws.data_ptr(), ws.numel()
    if done is None:
        p.flags |= _lib.FLAG_BWD_EMITTED
        return 4                                            # (the caller returns ws to the free list after the launch)
    torch.cuda.current_stream(device).wait_event(done)
    p.flags |= _lib.FLAG_BWD_PRESORTED
    return 2


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the cross-view sampling path has no CPU "
                           f"fallback (the CPU oracle lives in oracle/ and is test-only)")


def _f32c(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------------------
# feature packing (NCHW -> channel-last), once per forward, shared by the 6 layers
# --------------------------------------------------------------------------------------
def _pack_raw(feat: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """(B,N,C,H,W) -> detached channel-last (B*N,H,W,C); zero-copy when the map is
    already channels_last in memory (a backbone run in torch.channels_last)."""
    B, N, Cc, H, W = feat.shape
    feat = feat.detach()
    if dtype == feat.dtype:
        try:
            nhwc = feat.view(B * N, Cc, H, W).permute(0, 2, 3, 1)
        except RuntimeError:
            nhwc = None
        if nhwc is not None and nhwc.is_contiguous():
            return nhwc
    src = feat.contiguous()
    dst = torch.empty((B * N, H, W, Cc), device=feat.device, dtype=dtype)
    st = _lib.load().gd4d_pack_nchw(src.data_ptr(), dst.data_ptr(), _dtype_code(src), _dtype_code(dst),
                                    B * N, Cc, H, W, _stream_ptr(feat.device))
    _lib.check(st, "gd4d_pack_nchw")
    _count()
    return dst


class _AttachFn(torch.autograd.Function):
    """Differentiable alias of an already packed level (dense value_proj path): the
    backward hands the channel-last gradient back as a (B,N,C,H,W) VIEW, no copy."""

    @staticmethod
    def forward(ctx, feat: torch.Tensor, packed: torch.Tensor):
        ctx.shape = tuple(feat.shape)
        ctx.in_dtype = feat.dtype
        return packed.view_as(packed)

    @staticmethod
    def backward(ctx, grad):
        B, N = ctx.shape[:2]
        g = grad.permute(0, 3, 1, 2).unflatten(0, (B, N))
        return (g if g.dtype == ctx.in_dtype else g.to(ctx.in_dtype)), None


class GradSink:
    """ONE fp32 channel-last gradient map per level, shared by every decoder layer of
    a forward: each layer's backward kernel accumulates into it with vector
    atomics, so the dense zero-fill and the dense read-modify-write happen once per
    step instead of once per layer (bytes_bwd)."""

    def __init__(self, levels: Sequence[torch.Tensor]):
        self._shapes = [tuple(v.shape) for v in levels]
        self._device = levels[0].device
        self.buffers: Optional[List[torch.Tensor]] = None

    def get(self) -> List[torch.Tensor]:
        if self.buffers is None:
            # one allocation, one memset for all levels
            sizes = [int(np.prod(s)) for s in self._shapes]
            flat = torch.zeros(sum(sizes), device=self._device, dtype=torch.float32)
            self.buffers, o = [], 0
            for s, n in zip(self._shapes, sizes):
                self.buffers.append(flat[o:o + n].view(s))
                o += n
        return self.buffers

    def take(self):
        b, self.buffers = self.buffers, None
        return b


class _SinkFn(torch.autograd.Function):
    """Graph node that owns the GradSink: its `token` output is an input of every
    layer's sampling op, so autograd runs this backward only after ALL layers have
    accumulated; it then returns the shared maps as (B,N,C,H,W) views."""

    @staticmethod
    def forward(ctx, sink: GradSink, *feats: torch.Tensor):
        ctx.sink = sink
        ctx.meta = [(tuple(f.shape), f.dtype, f.is_contiguous()) for f in feats]
        return feats[0].new_zeros(())

    @staticmethod
    def backward(ctx, _gtoken):
        bufs = ctx.sink.take()
        if bufs is None:
            return (None,) + (None,) * len(ctx.meta)
        outs = []
        for buf, (shape, dtype, nchw) in zip(bufs, ctx.meta):
            if nchw and dtype in (torch.float32, torch.bfloat16):
                # NCHW producer: one tiled transpose (gd4d_unpack_nhwc[_cast]) instead of the strided
                # elementwise copy autograd would make to match the leaf's / conv's layout; a bf16 producer
                # gets its bf16 gradient from the same launch (no separate 4-byte -> 2-byte pass)
                g = torch.empty(shape, device=buf.device, dtype=dtype)
                st = _lib.load().gd4d_unpack_nhwc_cast(buf.data_ptr(), g.data_ptr(), F32 if dtype == torch.float32 else BF16,
                                                       shape[0] * shape[1], shape[2], shape[3], shape[4],
                                                       _stream_ptr(buf.device))
                _lib.check(st, "gd4d_unpack_nhwc_cast")
                _count()
            else:                                          # channels-last producer: a view, no copy
                g = buf.permute(0, 3, 1, 2).unflatten(0, shape[:2])
                g = g if g.dtype == dtype else g.to(dtype)
            outs.append(g)
        return (None, *outs)


@dataclass
class PackedFeatures:
    """Channel-last feature maps of one forward, plus the batch/camera split."""
    levels: List[torch.Tensor]            # detached, each (B*N, H_l, W_l, C)
    B: int
    N: int
    token: Optional[torch.Tensor] = None  # autograd handle of the shared GradSink
    sink: Optional[GradSink] = None
    sources: Optional[List[torch.Tensor]] = field(default=None, repr=False)

    @property
    def C(self) -> int:
        return self.levels[0].shape[-1]

    @property
    def shapes(self) -> List[Tuple[int, int]]:
        return [(int(v.shape[1]), int(v.shape[2])) for v in self.levels]

    def differentiable_levels(self) -> List[torch.Tensor]:
        """Packed levels as autograd-connected tensors (dense value_proj path)."""
        if self.sources is None:
            return self.levels
        return [_AttachFn.apply(f, v) if f.requires_grad else v for f, v in zip(self.sources, self.levels)]


def pack_level(feat: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    _require_cuda(feat, "feature map")
    if feat.dim() != 5:
        raise ValueError(f"feature map must be (B,N,C,H,W), got {tuple(feat.shape)}")
    return _pack_raw(feat, dtype or feat.dtype)


def pack_features(mlvl_feats: Sequence[torch.Tensor], dtype: Optional[torch.dtype] = None) -> PackedFeatures:
    """Pack the reference's ``value`` (python list of (B,N,C,H,W) maps,
    detr3d_transformer.py:142-143) once; all decoder layers then share it."""
    B, N = int(mlvl_feats[0].shape[0]), int(mlvl_feats[0].shape[1])
    levels = [pack_level(f, dtype) for f in mlvl_feats]
    token = sink = None
    if torch.is_grad_enabled() and any(f.requires_grad for f in mlvl_feats):
        sink = GradSink(levels)
        token = _SinkFn.apply(sink, *mlvl_feats)
    return PackedFeatures(levels, B, N, token, sink, list(mlvl_feats))


def lidar2img_to_tensor(img_metas, device) -> torch.Tensor:
    """img_metas[b]['lidar2img'] (list of N np(4,4), float64 or float32) -> (B,N,4,4)
    fp32 on the device; float64->float32 rounding as reference_points.new_tensor
    does it (detr3d_transformer.py:398-402)."""
    arr = np.asarray([m["lidar2img"] for m in img_metas])
    return torch.as_tensor(arr.astype(np.float32)).to(device, non_blocking=True)


# --------------------------------------------------------------------------------------
# raw C-ABI calls
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class XViewConfig:
    mode: int
    num_heads: int
    num_points: int
    pc_range: Tuple[float, ...]
    img_h: float
    img_w: float
    wide: bool = False          # mode C: gather-then-project (include/gd4d_xview.h)


@dataclass(frozen=True)
class GenLayout:
    """Column offsets (in floats) of the three generator blocks inside one packed row-major
    (B*Q, width) matrix -- the output of ONE GEMM over the concatenated weights of
    cam_attention_weights / deform_sampling_offsets / attention_weights (gen_stride in
    include/gd4d_xview.h)."""
    cam: int
    offsets: int
    attn: int
    width: int


def _fill_params(cfg: XViewConfig, values: Sequence[torch.Tensor], B: int, N: int, ref, attn_logits,
                 offsets, cam_logits, lidar2img, gen=None, layout: Optional[GenLayout] = None) -> XViewParams:
    L = len(values)
    if L > _lib.MAX_LEVELS:
        raise ValueError(f"at most {_lib.MAX_LEVELS} feature levels")
    Cc = int(values[0].shape[-1])
    p = XViewParams()
    p.abi_version = _lib.ABI_VERSION
    p.mode = cfg.mode
    p.wide = 1 if cfg.wide else 0
    p.value_dtype = _dtype_code(values[0])
    p.B, p.Q, p.N = B, int(ref.shape[1]), N
    p.Hh = cfg.num_heads if cfg.mode != MODE_A else Cc // 32
    p.L, p.P, p.C = L, cfg.num_points, Cc
    for l, v in enumerate(values):
        if v.dim() != 4 or v.shape[0] != B * N or v.shape[-1] != Cc or not v.is_contiguous():
            raise ValueError("value levels must be contiguous channel-last (B*N,H,W,C)")
        if v.dtype != values[0].dtype or v.device != values[0].device:
            raise ValueError("all value levels must share dtype and device")
        p.level_h[l], p.level_w[l] = int(v.shape[1]), int(v.shape[2])
        p.value[l] = v.data_ptr()
    pc = cfg.pc_range
    for i in range(3):
        p.pc_lo[i] = float(pc[i])
        p.pc_span[i] = float(pc[3 + i] - pc[i])      # python double subtraction, then fp32
    p.img_h, p.img_w = float(cfg.img_h), float(cfg.img_w)
    p.ref = ref.data_ptr()
    p.lidar2img = lidar2img.data_ptr()
    if gen is not None:
        base = gen.data_ptr()
        p.attn_logits, p.offsets, p.cam_logits = base + 4 * layout.attn, base + 4 * layout.offsets, base + 4 * layout.cam
        p.gen_stride = layout.width
    else:
        p.attn_logits = attn_logits.data_ptr()
        p.offsets = offsets.data_ptr() if offsets is not None else None
        p.cam_logits = cam_logits.data_ptr() if cam_logits is not None else None
    if DYNAMIC_SCHEDULE and cfg.mode == MODE_C:
        # mode A launches are ~10 us of work: the per-item claim costs more than the tail it removes
        p.sched = _sched_ptr(ref.device)
        if TMA_FORWARD and cfg.wide:
            p.flags = _lib.FLAG_TMA_FORWARD
    if L2_PREFETCH:
        p.flags |= _lib.FLAG_L2_PREFETCH
    return p


def _check_shapes(cfg, B, N, L, ref, attn_logits, offsets, cam_logits, lidar2img):
    Q = ref.shape[1]
    if tuple(ref.shape) != (B, Q, 3):
        raise ValueError(f"reference_points must be (B,Q,3), got {tuple(ref.shape)}")
    if tuple(lidar2img.shape) != (B, N, 4, 4):
        raise ValueError(f"lidar2img must be (B,N,4,4)=({B},{N},4,4), got {tuple(lidar2img.shape)}")
    P = cfg.num_points
    if cfg.mode == MODE_A:
        want = B * Q * N * P * L
    elif cfg.mode == MODE_V2:
        want = B * Q * N * cfg.num_heads * L * P
        if offsets is None or offsets.numel() != want * 2:
            raise ValueError("V2 offsets must hold B*Q*N*Hh*L*P*2 values")
        if L != P:
            raise ValueError("Detr3DCrossAttenV2 semantics need num_points == num_levels "
                             "(detr3d_transformer.py:611 vs :709)")
    else:
        want = B * Q * cfg.num_heads * L * P
        if offsets is None or offsets.numel() != B * Q * cfg.num_heads * P * 3:
            raise ValueError("offsets must hold B*Q*Hh*P*3 values")
        if cam_logits is None or cam_logits.numel() != B * Q * N:
            raise ValueError("cam_logits must hold B*Q*N values")
    if attn_logits.numel() != want:
        raise ValueError(f"attn_logits holds {attn_logits.numel()} values, expected {want}")


def _out_shape(cfg, B, Q, Cc):
    return (B, cfg.num_heads, Q, Cc) if cfg.wide else (B, Q, Cc)


def xview_forward(cfg: XViewConfig, values: Sequence[torch.Tensor], B: int, N: int, ref, attn_logits,
                  offsets=None, cam_logits=None, lidar2img=None, want_mask: bool = False, emit_for_backward: bool = False):
    """One fused forward launch.
    narrow: returns (out (B,Q,C), mask|None);  wide: returns ((out (B,Hh,Q,C), wsum (B,Hh,Q)), mask|None)."""
    for v in values:
        _require_cuda(v, "value")
    ref = _f32c(ref, "reference_points")
    attn_logits = _f32c(attn_logits, "attn_logits")
    offsets = _f32c(offsets, "offsets")
    cam_logits = _f32c(cam_logits, "cam_logits")
    lidar2img = _f32c(lidar2img, "lidar2img")
    L, Cc = len(values), int(values[0].shape[-1])
    _check_shapes(cfg, B, N, L, ref, attn_logits, offsets, cam_logits, lidar2img)
    p = _fill_params(cfg, values, B, N, ref, attn_logits, offsets, cam_logits, lidar2img)
    Q = int(ref.shape[1])
    out = torch.empty(_out_shape(cfg, B, Q, Cc), device=ref.device, dtype=torch.float32)
    p.out = out.data_ptr()
    wsum = None
    if cfg.wide:
        wsum = torch.empty((B, cfg.num_heads, Q), device=ref.device, dtype=torch.float32)
        p.wsum = wsum.data_ptr()
    mask = None
    if want_mask:
        shape = (B, Q, N) if cfg.mode != MODE_C else (B, N, Q, cfg.num_heads, cfg.num_points)
        mask = torch.zeros(shape, device=ref.device, dtype=torch.uint8)
        p.mask = mask.data_ptr()
    emitted = _emit_scratch(p, ref.device, emit_for_backward)
    st = _lib.load().gd4d_xview_forward(C.byref(p), _stream_ptr(ref.device))
    _lib.check(st, "gd4d_xview_forward")
    _count()
    if emit_for_backward:
        return ((out, wsum) if cfg.wide else out), mask, emitted
    return ((out, wsum) if cfg.wide else out), mask


def xview_backward(cfg: XViewConfig, values: Sequence[torch.Tensor], B: int, N: int, ref, attn_logits,
                   offsets, cam_logits, lidar2img, grad_out, grad_values: Optional[Sequence[torch.Tensor]],
                   need_ref: bool = True, need_offsets: bool = True, grad_wsum=None, presort=None):
    """One fused backward launch.  ``grad_values`` (fp32 channel-last, same shapes as
    ``values``) are ACCUMULATED into; the small gradients are returned fresh."""
    grad_out = _f32c(grad_out, "grad_out")
    p = _fill_params(cfg, values, B, N, ref, attn_logits, offsets, cam_logits, lidar2img)
    p.grad_out = grad_out.data_ptr()
    if grad_wsum is not None:
        grad_wsum = _f32c(grad_wsum, "grad_wsum")
        p.grad_wsum = grad_wsum.data_ptr()
    if grad_values is not None:
        for l, gvl in enumerate(grad_values):
            if gvl.dtype != torch.float32 or gvl.shape != values[l].shape or not gvl.is_contiguous():
                raise ValueError("grad_values must be fp32, contiguous, shaped like the value levels")
            p.grad_value[l] = gvl.data_ptr()
    # one zero-fill for all small gradients
    n_attn = attn_logits.numel()
    n_off = offsets.numel() if (offsets is not None and need_offsets) else 0
    n_cam = cam_logits.numel() if cam_logits is not None else 0
    n_ref = ref.numel() if need_ref else 0
    small = torch.zeros(n_attn + n_off + n_cam + n_ref, device=ref.device, dtype=torch.float32)
    g_attn = small[:n_attn].view(attn_logits.shape)
    o = n_attn
    g_off = small[o:o + n_off].view(offsets.shape) if n_off else None
    o += n_off
    g_cam = small[o:o + n_cam].view(cam_logits.shape) if n_cam else None
    o += n_cam
    g_ref = small[o:o + n_ref].view(ref.shape) if n_ref else None
    p.grad_attn_logits = g_attn.data_ptr()
    p.grad_offsets = g_off.data_ptr() if g_off is not None else None
    p.grad_cam_logits = g_cam.data_ptr() if g_cam is not None else None
    p.grad_ref = g_ref.data_ptr() if g_ref is not None else None
    n_launch = _use_presorted(p, presort, ref.device) if presort is not None else _attach_bwd_ws(p, ref.device)
    st = _lib.load().gd4d_xview_backward(C.byref(p), _stream_ptr(ref.device))
    _lib.check(st, "gd4d_xview_backward")
    _count(n_launch)
    _recycle_scratch(presort)
    return g_attn, g_off, g_cam, g_ref


class PreparedLaunch:
    """A fully-filled parameter block that can be re-launched without any host-side
    tensor bookkeeping (kernel-only timing loops, CUDA-graph capture)."""

    def __init__(self, fn, what, params, device, keepalive, kernels=1):
        self._fn, self._what, self.params, self._device, self._keep = fn, what, params, device, keepalive
        self.kernels = kernels

    def launch(self):
        st = self._fn(C.byref(self.params), _stream_ptr(self._device))
        if st != 0:
            _lib.check(st, self._what)
        _count(self.kernels)


def prepare_forward(cfg, values, B, N, ref, attn_logits, offsets, cam_logits, lidar2img) -> PreparedLaunch:
    ref, attn_logits, lidar2img = _f32c(ref, "ref"), _f32c(attn_logits, "attn"), _f32c(lidar2img, "l2i")
    offsets, cam_logits = _f32c(offsets, "offsets"), _f32c(cam_logits, "cam")
    _check_shapes(cfg, B, N, len(values), ref, attn_logits, offsets, cam_logits, lidar2img)
    p = _fill_params(cfg, values, B, N, ref, attn_logits, offsets, cam_logits, lidar2img)
    Q, Cc = int(ref.shape[1]), int(values[0].shape[-1])
    out = torch.empty(_out_shape(cfg, B, Q, Cc), device=ref.device, dtype=torch.float32)
    p.out = out.data_ptr()
    wsum = None
    if cfg.wide:
        wsum = torch.empty((B, cfg.num_heads, Q), device=ref.device, dtype=torch.float32)
        p.wsum = wsum.data_ptr()
    pl = PreparedLaunch(_lib.load().gd4d_xview_forward, "gd4d_xview_forward", p, ref.device,
                        (list(values), ref, attn_logits, offsets, cam_logits, lidar2img, out, wsum))
    pl.out, pl.wsum = out, wsum
    return pl


def prepare_backward(cfg, values, B, N, ref, attn_logits, offsets, cam_logits, lidar2img, grad_out,
                     grad_values, grad_wsum=None) -> PreparedLaunch:
    """Backward launch that keeps accumulating into ``grad_values`` and one small-gradient
    buffer (never re-zeroed: for timing only)."""
    ref, attn_logits, lidar2img = _f32c(ref, "ref"), _f32c(attn_logits, "attn"), _f32c(lidar2img, "l2i")
    offsets, cam_logits, grad_out = _f32c(offsets, "offsets"), _f32c(cam_logits, "cam"), _f32c(grad_out, "gout")
    p = _fill_params(cfg, values, B, N, ref, attn_logits, offsets, cam_logits, lidar2img)
    p.grad_out = grad_out.data_ptr()
    grad_wsum = _f32c(grad_wsum, "grad_wsum")
    if grad_wsum is not None:
        p.grad_wsum = grad_wsum.data_ptr()
    for l, gvl in enumerate(grad_values):
        p.grad_value[l] = gvl.data_ptr()
    smalls = [torch.zeros_like(t) for t in (attn_logits, ref)]
    p.grad_attn_logits, p.grad_ref = smalls[0].data_ptr(), smalls[1].data_ptr()
    if offsets is not None:
        smalls.append(torch.zeros_like(offsets))
        p.grad_offsets = smalls[-1].data_ptr()
    if cam_logits is not None:
        smalls.append(torch.zeros_like(cam_logits))
        p.grad_cam_logits = smalls[-1].data_ptr()
    n_launch = _attach_bwd_ws(p, ref.device)
    return PreparedLaunch(_lib.load().gd4d_xview_backward, "gd4d_xview_backward", p, ref.device,
                          (list(values), ref, attn_logits, offsets, cam_logits, lidar2img, grad_out,
                           grad_wsum, list(grad_values), smalls), kernels=n_launch)


def _check_gen(cfg, B, N, L, ref, gen, layout: GenLayout, lidar2img):
    Q = ref.shape[1]
    if cfg.mode != MODE_C:
        raise ValueError("the packed generator layout is a mode C feature")
    if tuple(ref.shape) != (B, Q, 3) or tuple(lidar2img.shape) != (B, N, 4, 4):
        raise ValueError("bad reference_points / lidar2img shape")
    if tuple(gen.shape) != (B, Q, layout.width):
        raise ValueError(f"packed generator output must be (B,Q,{layout.width}), got {tuple(gen.shape)}")
    blocks = sorted([(layout.cam, N), (layout.offsets, cfg.num_heads * cfg.num_points * 3),
                     (layout.attn, cfg.num_heads * L * cfg.num_points)])
    end = 0
    for start, width in blocks:
        if start < end:
            raise ValueError("generator blocks overlap")
        end = start + width
    if end > layout.width:
        raise ValueError("generator blocks exceed the packed width")


def xview_forward_gen(cfg: XViewConfig, values, B: int, N: int, ref, gen, layout: GenLayout, lidar2img,
                      emit_for_backward: bool = False):
    """Mode C forward reading logits / offsets / camera logits as column blocks of ``gen``."""
    for v in values:
        _require_cuda(v, "value")
    ref, gen, lidar2img = _f32c(ref, "reference_points"), _f32c(gen, "gen"), _f32c(lidar2img, "lidar2img")
    L, Cc = len(values), int(values[0].shape[-1])
    _check_gen(cfg, B, N, L, ref, gen, layout, lidar2img)
    p = _fill_params(cfg, values, B, N, ref, None, None, None, lidar2img, gen=gen, layout=layout)
    Q = int(ref.shape[1])
    out = torch.empty(_out_shape(cfg, B, Q, Cc), device=ref.device, dtype=torch.float32)
    p.out = out.data_ptr()
    wsum = None
    if cfg.wide:
        wsum = torch.empty((B, cfg.num_heads, Q), device=ref.device, dtype=torch.float32)
        p.wsum = wsum.data_ptr()
    emitted = _emit_scratch(p, ref.device, emit_for_backward)
    _lib.check(_lib.load().gd4d_xview_forward(C.byref(p), _stream_ptr(ref.device)), "gd4d_xview_forward")
    _count()
    if emit_for_backward:
        return ((out, wsum) if cfg.wide else out), emitted
    return (out, wsum) if cfg.wide else out


def xview_backward_gen(cfg: XViewConfig, values, B: int, N: int, ref, gen, layout: GenLayout, lidar2img,
                       grad_out, grad_values, need_ref: bool = True, grad_wsum=None, presort=None):
    """Backward of ``xview_forward_gen``: returns (grad_gen (B,Q,width), grad_ref|None), one
    zero-filled allocation for both."""
    grad_out = _f32c(grad_out, "grad_out")
    p = _fill_params(cfg, values, B, N, ref, None, None, None, lidar2img, gen=gen, layout=layout)
    p.grad_out = grad_out.data_ptr()
    if grad_wsum is not None:
        grad_wsum = _f32c(grad_wsum, "grad_wsum")
        p.grad_wsum = grad_wsum.data_ptr()
    if grad_values is not None:
        for l, gvl in enumerate(grad_values):
            if gvl.dtype != torch.float32 or gvl.shape != values[l].shape or not gvl.is_contiguous():
                raise ValueError("grad_values must be fp32, contiguous, shaped like the value levels")
            p.grad_value[l] = gvl.data_ptr()
    n_gen, n_ref = gen.numel(), (ref.numel() if need_ref else 0)
    small = torch.zeros(n_gen + n_ref, device=ref.device, dtype=torch.float32)
    g_gen = small[:n_gen].view(gen.shape)
    g_ref = small[n_gen:].view(ref.shape) if need_ref else None
    base = g_gen.data_ptr()
    p.grad_attn_logits = base + 4 * layout.attn
    p.grad_offsets = base + 4 * layout.offsets
    p.grad_cam_logits = base + 4 * layout.cam
    p.grad_ref = g_ref.data_ptr() if g_ref is not None else None
    n_launch = _use_presorted(p, presort, ref.device) if presort is not None else _attach_bwd_ws(p, ref.device)
    _lib.check(_lib.load().gd4d_xview_backward(C.byref(p), _stream_ptr(ref.device)), "gd4d_xview_backward")
    _count(n_launch)
    _recycle_scratch(presort)
    return g_gen, g_ref


# --------------------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------------------
_I_REF, _I_ATTN, _I_OFF, _I_CAM, _I_TOKEN, _I_VALUES = 3, 4, 5, 6, 8, 10


class _XViewFn(torch.autograd.Function):
    """inputs: cfg, B, N, ref, attn_logits, offsets, cam_logits, lidar2img, token, sink, *values

    With a ``token``/``sink`` pair (shared, detached feature maps) the feature
    gradient is accumulated into the sink and only a dummy gradient flows to the
    token; otherwise ``values`` are differentiable inputs and get fresh fp32 grads.
    """

    @staticmethod
    def forward(ctx, cfg: XViewConfig, B: int, N: int, ref, attn_logits, offsets, cam_logits, lidar2img,
                token, sink, *values):
        ref_c = _f32c(ref, "reference_points")
        attn_c = _f32c(attn_logits, "attn_logits")
        off_c = _f32c(offsets, "offsets")
        cam_c = _f32c(cam_logits, "cam_logits")
        l2i_c = _f32c(lidar2img, "lidar2img")
        want = bool(cfg.wide and cfg.mode == MODE_C and any(ctx.needs_input_grad))
        r = xview_forward(cfg, values, B, N, ref_c, attn_c, off_c, cam_c, l2i_c, emit_for_backward=want)
        res, emitted = (r[0], r[2]) if want else (r[0], None)
        ctx.cfg, ctx.B, ctx.N, ctx.sink = cfg, B, N, sink
        ctx.has_off, ctx.has_cam = offsets is not None, cam_logits is not None
        empty = ref_c.new_empty(0)
        ctx.save_for_backward(ref_c, attn_c, off_c if off_c is not None else empty,
                              cam_c if cam_c is not None else empty, l2i_c, *values)
        ctx.presort = emitted
        if emitted is None and cfg.wide and cfg.mode == MODE_C and any(ctx.needs_input_grad):
            ctx.presort = _presort(_fill_params(cfg, values, B, N, ref_c, attn_c, off_c, cam_c, l2i_c), ref_c.device)
        return res

    @staticmethod
    def backward(ctx, grad_out, grad_wsum=None):
        ref, attn, off, cam, l2i, *values = ctx.saved_tensors
        off = off if ctx.has_off else None
        cam = cam if ctx.has_cam else None
        nd = ctx.needs_input_grad
        use_sink = ctx.sink is not None and nd[_I_TOKEN]
        need_values = any(nd[_I_VALUES:])
        grad_values = None
        if use_sink:
            grad_values = ctx.sink.get()
        elif need_values:
            grad_values = [torch.zeros(v.shape, device=v.device, dtype=torch.float32) for v in values]
        g_attn, g_off, g_cam, g_ref = xview_backward(
            ctx.cfg, values, ctx.B, ctx.N, ref, attn, off, cam, l2i, grad_out, grad_values,
            need_ref=nd[_I_REF], need_offsets=ctx.has_off and nd[_I_OFF],
            grad_wsum=grad_wsum if ctx.cfg.wide else None, presort=ctx.presort)
        ctx.presort = None
        gv = [None] * len(values)
        if need_values and not use_sink:
            gv = [g if g.dtype == v.dtype else g.to(v.dtype) for g, v in zip(grad_values, values)]
        g_token = grad_out.new_zeros(()) if use_sink else None
        return (None, None, None, g_ref if nd[_I_REF] else None, g_attn if nd[_I_ATTN] else None,
                g_off if (ctx.has_off and nd[_I_OFF]) else None,
                g_cam if (ctx.has_cam and nd[_I_CAM]) else None, None, g_token, None, *gv)


class _XViewGenFn(torch.autograd.Function):
    """inputs: cfg, B, N, ref, gen, layout, lidar2img, token, sink, *values  (mode C, packed
    generator output; same token/sink contract as _XViewFn)."""

    @staticmethod
    def forward(ctx, cfg: XViewConfig, B: int, N: int, ref, gen, layout: GenLayout, lidar2img,
                token, sink, *values):
        ref_c, gen_c, l2i_c = _f32c(ref, "reference_points"), _f32c(gen, "gen"), _f32c(lidar2img, "lidar2img")
        want = bool(cfg.wide and any(ctx.needs_input_grad))
        r = xview_forward_gen(cfg, values, B, N, ref_c, gen_c, layout, l2i_c, emit_for_backward=want)
        res, emitted = r if want else (r, None)
        ctx.cfg, ctx.B, ctx.N, ctx.sink, ctx.layout = cfg, B, N, sink, layout
        ctx.save_for_backward(ref_c, gen_c, l2i_c, *values)
        ctx.presort = emitted
        if emitted is None and cfg.wide and any(ctx.needs_input_grad):
            ctx.presort = _presort(_fill_params(cfg, values, B, N, ref_c, None, None, None, l2i_c, gen=gen_c,
                                                layout=layout), ref_c.device)
        return res

    @staticmethod
    def backward(ctx, grad_out, grad_wsum=None):
        ref, gen, l2i, *values = ctx.saved_tensors
        nd = ctx.needs_input_grad
        use_sink = ctx.sink is not None and nd[7]
        need_values = any(nd[9:])
        grad_values = None
        if use_sink:
            grad_values = ctx.sink.get()
        elif need_values:
            grad_values = [torch.zeros(v.shape, device=v.device, dtype=torch.float32) for v in values]
        g_gen, g_ref = xview_backward_gen(ctx.cfg, values, ctx.B, ctx.N, ref, gen, ctx.layout, l2i, grad_out,
                                          grad_values, need_ref=nd[3],
                                          grad_wsum=grad_wsum if ctx.cfg.wide else None, presort=ctx.presort)
        ctx.presort = None
        gv = [None] * len(values)
        if need_values and not use_sink:
            gv = [g if g.dtype == v.dtype else g.to(v.dtype) for g, v in zip(grad_values, values)]
        g_token = grad_out.new_zeros(()) if use_sink else None
        return (None, None, None, g_ref if nd[3] else None, g_gen if nd[4] else None, None, None, g_token,
                None, *gv)


def xview_attention_gen(cfg: XViewConfig, packed: PackedFeatures, ref, gen, layout: GenLayout, lidar2img,
                        values: Optional[Sequence[torch.Tensor]] = None):
    """``xview_attention`` for mode C with the three generator outputs packed in ONE
    (B,Q,layout.width) tensor (no split copies forward, no concatenation backward)."""
    if values is not None:
        return _XViewGenFn.apply(cfg, packed.B, packed.N, ref, gen, layout, lidar2img, None, None, *values)
    return _XViewGenFn.apply(cfg, packed.B, packed.N, ref, gen, layout, lidar2img, packed.token, packed.sink,
                             *packed.levels)


def xview_attention(cfg: XViewConfig, packed: PackedFeatures, ref, attn_logits, offsets=None,
                    cam_logits=None, lidar2img=None, values: Optional[Sequence[torch.Tensor]] = None):
    """Differentiable fused cross-view sampling attention.

    narrow -> (B,Q,C) fp32;  wide -> (out (B,Hh,Q,C), wsum (B,Hh,Q)): head-major, so the
    per-head value_proj slices apply as one strided-batched GEMM without a transpose copy.
    ``values`` (differentiable channel-last tensors, e.g. the value_proj'ed maps of the
    dense path) override ``packed.levels``; without them the shared packed maps are
    sampled and their gradient goes to ``packed.sink``."""
    if values is not None:
        return _XViewFn.apply(cfg, packed.B, packed.N, ref, attn_logits, offsets, cam_logits, lidar2img,
                              None, None, *values)
    return _XViewFn.apply(cfg, packed.B, packed.N, ref, attn_logits, offsets, cam_logits, lidar2img,
                          packed.token, packed.sink, *packed.levels)
