"""TEST INFRASTRUCTURE ONLY -- the reference's per-layer loss normalisers, restated line by line
(second half):

    Detr3DHead.loss_single   projects/mmdet3d_plugin/models/dense_heads/detr3d_head.py:316-331
    reduce_mean              mmdet 2.x mmdet/core/utils/dist_utils.py (third-party, un-vendored, un-pinned):
                                 if not (dist.is_available() and dist.is_initialized()): return tensor
                                 tensor = tensor.clone()
                                 dist.all_reduce(tensor.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
                                 return tensor

One call per decoder layer, each with its own one-element tensors, all-reduces and ``.item()``, exactly
like the reference.  Parity unpinned by reference fixtures (the reference has none for the loss); the
lines are few enough to read against the source.  Only tests/ may import this file.
"""
import torch
import torch.distributed as dist


def reduce_mean(tensor):
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    tensor = tensor.clone()
    dist.all_reduce(tensor.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
    return tensor


def layer_avg_factors(num_total_pos: int, num_total_neg: int, bg_cls_weight: float, sync_cls_avg_factor: bool,
                      like: torch.Tensor):
    """detr3d_head.py:316-331 for ONE layer -> (cls_avg_factor, num_total_pos) as the loss functions get them."""
    cls_avg_factor = num_total_pos * 1.0 + num_total_neg * bg_cls_weight            # :316-317
    if sync_cls_avg_factor:                                                         # :318-320
        cls_avg_factor = reduce_mean(like.new_tensor([cls_avg_factor]))
    cls_avg_factor = max(cls_avg_factor, 1)                                         # :322
    num_total_pos = like.new_tensor([num_total_pos])                                # :328
    num_total_pos = torch.clamp(reduce_mean(num_total_pos), min=1).item()           # :329
    return cls_avg_factor, num_total_pos
