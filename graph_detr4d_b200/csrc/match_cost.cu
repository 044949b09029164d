// Fused Hungarian match-cost kernel (include/gd4d_assign.h): the cost
// matrices of all decoder layers of one sample in one launch -- FocalLossCost (mmdet 2.x formula,
// gamma = 2) + BBox3DL1Cost over the first 8 box-code entries against normalize_bbox(gt), then
// nan_to_num, as hungarian_assigner_3d.py:117-131 computes them with ~15 torch ops per layer.
// One thread per (row, gt): the gt row is normalised on the fly (3 logs, sin, cos).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gd4d_assign.h"

namespace gd4d {

__global__ void __launch_bounds__(256)
match_cost_kernel(const float* __restrict__ cls_pred, const float* __restrict__ bbox_pred,
                  const float* __restrict__ gt, const int64_t* __restrict__ labels,
                  float* __restrict__ cost, int64_t rows, int C, int code, int G, int gt_dim,
                  float cls_w, float reg_w, float alpha, float eps) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * G) return;
  const int64_t row = i / G;
  const int g = static_cast<int>(i - row * G);
  const float* b = gt + static_cast<size_t>(g) * gt_dim;
  // normalize_bbox (util.py:38-57): (cx, cy, log w, log l, cz, log h, sin rot, cos rot)
  const float n[8] = {b[0], b[1], logf(b[3]), logf(b[4]), b[2], logf(b[5]), sinf(b[6]), cosf(b[6])};
  const float* p = bbox_pred + static_cast<size_t>(row) * code;
  float reg = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) reg += fabsf(p[k] - n[k]);                       // torch.cdist(p=1)
  const int64_t lab = labels[g];
  if (lab < 0 || lab >= C) {   // ignore / background label: never index cls_pred out of bounds; the host
    cost[i] = 100.f;           // side (assign.py) rejects such labels after its one sync
    return;
  }
  const float x = cls_pred[static_cast<size_t>(row) * C + lab];
  const float s = __fdiv_rn(1.f, 1.f + expf(-x));                              // sigmoid
  const float neg = -logf((1.f - s) + eps) * (1.f - alpha) * (s * s);
  const float om = 1.f - s;
  const float pos = -logf(s + eps) * alpha * (om * om);
  float c = (pos - neg) * cls_w + reg * reg_w;
  if (isnan(c)) c = 100.f;                                                     // nan_to_num (:131)
  else if (isinf(c)) c = c > 0.f ? 100.f : -100.f;
  cost[i] = c;
}

}  // namespace gd4d

extern "C" int gd4d_match_cost(const float* cls_pred, const float* bbox_pred, const float* gt_bboxes,
                               const int64_t* gt_labels, float* cost, int64_t rows, int32_t num_classes,
                               int32_t code_size, int32_t G, int32_t gt_dim, float cls_weight,
                               float reg_weight, float alpha, float eps, void* cuda_stream) {
  if (cls_pred == nullptr || bbox_pred == nullptr || gt_bboxes == nullptr || gt_labels == nullptr ||
      cost == nullptr)
    return GD4D_ERR_NULL;
  if (rows <= 0 || num_classes <= 0 || code_size < 8 || G <= 0 || gt_dim < 7) return GD4D_ERR_DIMS;
  const int64_t n = rows * G;
  if (n > (1LL << 40)) return GD4D_ERR_DIMS;
  gd4d::match_cost_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                            static_cast<cudaStream_t>(cuda_stream)>>>(
      cls_pred, bbox_pred, gt_bboxes, gt_labels, cost, rows, num_classes, code_size, G, gt_dim,
      cls_weight, reg_weight, alpha, eps);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}
