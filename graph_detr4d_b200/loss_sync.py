"""Loss normalisers of ALL decoder layers with one packed all-reduce and no host sync
(second half).

The reference's ``loss_single`` runs once per decoder layer
(projects/mmdet3d_plugin/models/dense_heads/detr3d_head.py:316-331; same lines in
detr3d_head_pe.py:822-837) and, every time,

    cls_avg_factor = num_total_pos * 1.0 + num_total_neg * self.bg_cls_weight
    if self.sync_cls_avg_factor:
        cls_avg_factor = reduce_mean(cls_scores.new_tensor([cls_avg_factor]))     # H2D + all-reduce
    cls_avg_factor = max(cls_avg_factor, 1)                                       # tensor > int: host sync
    ...
    num_total_pos = loss_cls.new_tensor([num_total_pos])                          # H2D
    num_total_pos = torch.clamp(reduce_mean(num_total_pos), min=1).item()         # all-reduce + host sync

i.e. 2 one-element host->device copies, 2 one-element all-reduces and 2 host syncs per layer: 12 + 12 + 12
per step for the 6-layer decoder.  At batch 1 those syncs, not NCCL bandwidth, cap data-parallel scaling
(SURVEY 8f).  The counts are host integers that are all known as soon as the (batched, one-sync) Hungarian
assignment returns (assign.BatchedHungarianAssigner3D: a Hungarian match has min(Q, G_b) positives), so here:

  * the 2L numbers travel in ONE pinned host->device copy,
  * are averaged over ranks by ONE all-reduce of that (2L,) tensor (mmdet's reduce_mean arithmetic:
    divide by the world size in fp32, then sum),
  * are clamped to >= 1 ON THE DEVICE and stay there: the loss functions divide by them as tensors
    (``loss.sum() / avg_factor``), which is the same fp32 division the reference performs with the
    ``.item()``-ed python float (up to torch's scalar-divisor shortcut on CUDA, which multiplies by the
    reciprocal: at most 1 ulp of the loss).  No host sync at all.

The arithmetic per element is the reference's, so the normalisers are bit-identical to its (checked
against the restated lines in oracle/loss_sync_oracle.py, single process and world size 2).
Host logic + torch.distributed plumbing: works on any device (the world-size-2 test runs it over gloo).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["packed_avg_factors", "hungarian_pos_neg_counts"]


def hungarian_pos_neg_counts(num_layers: int, num_query: int, gt_counts: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """(num_total_pos, num_total_neg) per decoder layer for a one-to-one Hungarian assignment of
    ``num_query`` predictions to ``gt_counts[b]`` ground-truth boxes per sample: every layer matches
    min(Q, G_b) queries of sample b (scipy's rectangular linear_sum_assignment), the rest are negatives
    (detr3d_head.py:270-271: the numel of pos_inds / neg_inds, summed over the images)."""
    pos = sum(min(int(num_query), int(g)) for g in gt_counts)
    neg = len(gt_counts) * int(num_query) - pos
    return np.full(num_layers, pos, dtype=np.int64), np.full(num_layers, neg, dtype=np.int64)


def packed_avg_factors(num_total_pos: Sequence[int], num_total_neg: Sequence[int], bg_cls_weight: float = 0.0,
                       sync_cls_avg_factor: bool = True, device=None,
                       group: Optional["dist.ProcessGroup"] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (cls_avg_factor (L,), num_total_pos (L,)) fp32 tensors on ``device``: what
    detr3d_head.py:316-331 computes layer by layer, for all L layers at once."""
    pos = np.asarray(num_total_pos, dtype=np.float64).reshape(-1)
    neg = np.asarray(num_total_neg, dtype=np.float64).reshape(-1)
    if pos.shape != neg.shape:
        raise ValueError("num_total_pos and num_total_neg must have one entry per decoder layer")
    L = pos.shape[0]
    cls = pos * 1.0 + neg * float(bg_cls_weight)                  # python-float arithmetic in the reference (:316-317)
    host = torch.from_numpy(np.concatenate([cls, pos]).astype(np.float32))        # new_tensor([...]): fp32
    device = torch.device(device) if device is not None else torch.device("cpu")
    if device.type == "cuda":
        host = host.pin_memory()
    buf = host.to(device, non_blocking=True)                       # ONE host -> device copy
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            # mmdet reduce_mean: tensor.clone().div_(world) then all_reduce(SUM)
            part = buf if sync_cls_avg_factor else buf[L:]
            part.div_(world)
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)              # ONE collective
    cls_avg_factor = buf[:L].clamp(min=1)                          # max(cls_avg_factor, 1)          (:322)
    num_pos = buf[L:].clamp(min=1)                                 # clamp(reduce_mean(.), min=1)    (:329)
    return cls_avg_factor, num_pos
