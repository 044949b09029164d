/* TEST INFRASTRUCTURE ONLY -- independent plain-C restatements used to cross-check the torch
 * oracles of the two "next" rows with nothing shared but the reference's formulas:
 *
 *   pe_frustum   Detr3DHeadPE.position_embeding, dense_heads/detr3d_head_pe.py:439-480
 *                (frustum point -> img2lidar -> normalise -> mask count -> inverse_sigmoid),
 *                output in the reference's (B*N, D*3, H, W) layout
 *   match_cost   HungarianAssigner3D cost matrix, core/bbox/assigners/hungarian_assigner_3d.py:117-131
 *                (mmdet 2.x FocalLossCost formula + BBox3DL1Cost on normalize_bbox(gt), nan_to_num)
 *
 * Build with -O2 -ffp-contract=off (oracle/c/Makefile).  Never linked into the product library.
 */
#include <math.h>
#include <stdint.h>

int pe_frustum(const float* img2lidar /* (BN,16) */, const uint8_t* mask_in /* (BN,H,W) or NULL */,
               float* out /* (BN, 3D, H, W) */, uint8_t* mask_out /* (BN,H,W) */, int BN, int H, int W, int D,
               float pad_h, float pad_w, float depth_start, float bin_size, const float* lo, const float* span) {
  const float eps = 1e-5f;
  for (int bn = 0; bn < BN; ++bn) {
    const float* M = img2lidar + (long)bn * 16;
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const float xw = ((float)w * pad_w) / (float)W;                       /* :440 */
        const float yh = ((float)h * pad_h) / (float)H;                       /* :439 */
        int outside = 0;
        for (int d = 0; d < D; ++d) {
          const float idx = (float)d;
          const float z = depth_start + (bin_size * idx) * (idx + 1.0f);      /* :455 */
          const float s = z > eps ? z : eps;                                  /* :460 */
          const float px = xw * s, py = yh * s;
          for (int r = 0; r < 3; ++r) {
            float acc = M[4 * r + 0] * px;                                    /* :468, sequential */
            acc = acc + M[4 * r + 1] * py;
            acc = acc + M[4 * r + 2] * z;
            acc = acc + M[4 * r + 3] * 1.0f;
            const float c = (acc - lo[r]) / span[r];                          /* :469-474 */
            outside += (c > 1.0f) || (c < 0.0f);                              /* :476 */
            float xc = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);               /* inverse_sigmoid */
            const float x1 = xc > eps ? xc : eps;
            const float om = 1.0f - xc;
            const float x2 = om > eps ? om : eps;
            out[(((long)bn * 3 * D + d * 3 + r) * H + h) * W + w] = logf(x1 / x2);
          }
        }
        const long mi = ((long)bn * H + h) * W + w;
        const int m = (float)outside > (float)D * 0.5f;                       /* :477 */
        mask_out[mi] = (uint8_t)(m || (mask_in != 0 && mask_in[mi] != 0));    /* :478 */
      }
  }
  return 0;
}

int match_cost(const float* cls_pred /* (Q,C) */, const float* bbox_pred /* (Q,code) */,
               const float* gt /* (G,gt_dim) */, const int64_t* labels, float* cost /* (Q,G) */, int Q, int C,
               int code, int G, int gt_dim, float cls_w, float reg_w, float alpha, float eps) {
  for (int q = 0; q < Q; ++q)
    for (int g = 0; g < G; ++g) {
      const float* b = gt + (long)g * gt_dim;
      const float n[8] = {b[0], b[1], logf(b[3]), logf(b[4]), b[2], logf(b[5]), sinf(b[6]), cosf(b[6])};
      float reg = 0.0f;
      for (int k = 0; k < 8; ++k) reg += fabsf(bbox_pred[(long)q * code + k] - n[k]);
      const float x = cls_pred[(long)q * C + labels[g]];
      const float s = 1.0f / (1.0f + expf(-x));
      const float neg = -logf((1.0f - s) + eps) * (1.0f - alpha) * (s * s);
      const float pos = -logf(s + eps) * alpha * ((1.0f - s) * (1.0f - s));
      float c = (pos - neg) * cls_w + reg * reg_w;
      if (isnan(c)) c = 100.0f;
      else if (isinf(c)) c = c > 0.0f ? 100.0f : -100.0f;
      cost[(long)q * G + g] = c;
    }
  return 0;
}
