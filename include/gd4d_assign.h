/*
 * gd4d_assign.h -- C ABI of the fused Hungarian match-cost kernel.
 *
 * The reference's loss assigns targets layer by layer (6 decoder layers x samples), each time
 * building the cost matrix with ~15 small torch ops, copying it to the host (a device sync) and
 * running scipy's linear_sum_assignment:
 *
 *   HungarianAssigner3D.assign   projects/mmdet3d_plugin/core/bbox/assigners/hungarian_assigner_3d.py:60-145
 *   BBox3DL1Cost                 projects/mmdet3d_plugin/core/bbox/match_costs/match_cost.py:6-28
 *   normalize_bbox               projects/mmdet3d_plugin/core/bbox/util.py:38-57
 *   FocalLossCost                mmdet 2.x (third-party, un-vendored; detr3d_res50.py:112)
 *
 * gd4d_match_cost computes the weighted cost matrices of ALL layers of one sample in one launch,
 * straight into a caller-provided slice of one buffer, so the whole step needs ONE device-to-host
 * copy and ONE sync (graph_detr4d_b200/assign.py).  fp32, the reference's op order:
 *
 *   p = sigmoid(cls_pred[row, label_g])
 *   cls = ( -log(p + eps)*alpha*(1-p)^2  -  -log(1 - p + eps)*(1-alpha)*p^2 ) * cls_weight
 *   reg = sum_{k<8} | bbox_pred[row,k] - normalize_bbox(gt_g)[k] |  * reg_weight
 *   cost[row,g] = nan_to_num(cls + reg, nan=100, posinf=100, neginf=-100)
 */
#ifndef GD4D_ASSIGN_H_
#define GD4D_ASSIGN_H_

#include "gd4d_xview.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cls_pred (rows, num_classes) logits; bbox_pred (rows, code_size >= 8); gt_bboxes (G, gt_dim >= 7)
 * un-normalised (cx, cy, cz, w, l, h, rot, ...); gt_labels (G) int64; cost (rows, G) fp32.
 * rows = layers * queries of one sample.  All DEVICE pointers. */
GD4D_API int gd4d_match_cost(const float* cls_pred, const float* bbox_pred, const float* gt_bboxes,
                             const int64_t* gt_labels, float* cost, int64_t rows,
                             int32_t num_classes, int32_t code_size, int32_t G, int32_t gt_dim,
                             float cls_weight, float reg_weight, float alpha, float eps,
                             void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GD4D_ASSIGN_H_ */
