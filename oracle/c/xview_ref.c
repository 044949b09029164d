/* TEST INFRASTRUCTURE ONLY -- independent plain-C restatement of the forward of the
 * cross-view sampling attention, used to cross-check the torch oracle
 * (oracle/xview_oracle.py) with nothing shared but the reference's formulas:
 *
 *   mode 0 (A): detr3d_transformer.py:373-383, 397-438   (Detr3DCrossAtten + feature_sampling)
 *   mode 1 (C): deform3d_cross_attn.py:211-258, 274, 281-284, 320-324 + mmcv
 *               multi_scale_deformable_attn_pytorch (grid_sample formulation)
 *
 * Feature maps are read in the REFERENCE's own layout, NCHW per level: (B, N, C, H_l, W_l).
 * Build with -O2 -ffp-contract=off (no FMA contraction: the projection must round exactly
 * like the reference's sequential fp32 mat-vec) -- see oracle/c/Makefile.
 * Never linked into the product library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 8

typedef struct {
  int mode;                 /* 0 = A, 1 = C */
  int B, Q, N, Hh, L, P, C;
  int level_h[MAXL], level_w[MAXL];
  const float* value[MAXL]; /* (B,N,C,H,W) */
  const float* ref;         /* (B,Q,3) */
  const float* lidar2img;   /* (B,N,16) */
  const float* attn_logits; /* A: (B,Q,N,P,L)   C: (B,Q,Hh,L,P) */
  const float* offsets;     /* C: (B,Q,Hh,P,3) */
  const float* cam_logits;  /* C: (B,Q*N) viewed (B,N,Q) */
  float pc_lo[3], pc_span[3];
  float img_h, img_w;
  float* out;               /* (B,Q,C) */
  uint8_t* mask;            /* A: (B,Q,N)  C: (B,N,Q,Hh,P) */
} xref_params;

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* returns valid; u,v = normalised image coordinates */
static int project(const float* M, float X, float Y, float Z, float img_w, float img_h, int mode_c,
                   float* u, float* v) {
  volatile float cx = ((M[0] * X + M[1] * Y) + M[2] * Z) + M[3];
  volatile float cy = ((M[4] * X + M[5] * Y) + M[6] * Z) + M[7];
  volatile float cz = ((M[8] * X + M[9] * Y) + M[10] * Z) + M[11];
  const float eps = 1e-5f;
  int ok = cz > eps;
  float den = cz > eps ? cz : eps;
  *u = (cx / den) / img_w;
  *v = (cy / den) / img_h;
  if (mode_c) {
    ok = ok && (*u > 0.f) && (*u < 1.f) && (*v > 0.f) && (*v < 1.f);
  } else {
    float gx = (*u - 0.5f) * 2.f, gy = (*v - 0.5f) * 2.f;
    ok = ok && (gx > -1.f) && (gx < 1.f) && (gy > -1.f) && (gy < 1.f);
  }
  return ok;
}

/* bilinear sample of channel c of one NCHW image, zeros padding, align_corners=False */
static float bilinear(const float* img, int c, int H, int W, float ix, float iy) {
  float fx = floorf(ix), fy = floorf(iy);
  int x0 = (int)fx, y0 = (int)fy;
  float tx = ix - fx, ty = iy - fy;
  const float* ch = img + (size_t)c * H * W;
  float acc = 0.f;
  if (y0 >= 0 && y0 < H) {
    if (x0 >= 0 && x0 < W) acc += ch[y0 * W + x0] * (1.f - tx) * (1.f - ty);
    if (x0 + 1 >= 0 && x0 + 1 < W) acc += ch[y0 * W + x0 + 1] * tx * (1.f - ty);
  }
  if (y0 + 1 >= 0 && y0 + 1 < H) {
    if (x0 >= 0 && x0 < W) acc += ch[(y0 + 1) * W + x0] * (1.f - tx) * ty;
    if (x0 + 1 >= 0 && x0 + 1 < W) acc += ch[(y0 + 1) * W + x0 + 1] * tx * ty;
  }
  return acc;
}

int xref_forward(const xref_params* p) {
  if (!p || p->L > MAXL) return -1;
  const int B = p->B, Q = p->Q, N = p->N, Hh = p->Hh, L = p->L, P = p->P, C = p->C;
  const int Ch = C / Hh;
  const int LP = L * P;
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
  for (int b = 0; b < B; ++b) {
    for (int q = 0; q < Q; ++q) {
      const float* r = p->ref + ((size_t)b * Q + q) * 3;
      const float X0 = r[0] * p->pc_span[0] + p->pc_lo[0];
      const float Y0 = r[1] * p->pc_span[1] + p->pc_lo[1];
      const float Z0 = r[2] * p->pc_span[2] + p->pc_lo[2];
      float* out = p->out + ((size_t)b * Q + q) * C;
      for (int c = 0; c < C; ++c) out[c] = 0.f;
      if (p->mode == 0) {
        for (int n = 0; n < N; ++n) {
          float u, v;
          int ok = project(p->lidar2img + ((size_t)b * N + n) * 16, X0, Y0, Z0, p->img_w, p->img_h, 0, &u, &v);
          if (p->mask) p->mask[((size_t)b * Q + q) * N + n] = (uint8_t)ok;
          if (!ok) continue;
          float gx = (u - 0.5f) * 2.f, gy = (v - 0.5f) * 2.f;
          for (int l = 0; l < L; ++l) {
            const int H = p->level_h[l], W = p->level_w[l];
            float wt = 0.f;
            for (int pp = 0; pp < P; ++pp)
              wt += sigmoidf_(p->attn_logits[((((size_t)b * Q + q) * N + n) * P + pp) * L + l]);
            float ix = (gx + 1.f) * (W * 0.5f) - 0.5f, iy = (gy + 1.f) * (H * 0.5f) - 0.5f;
            const float* img = p->value[l] + ((size_t)b * N + n) * C * H * W;
            for (int c = 0; c < C; ++c) out[c] += wt * bilinear(img, c, H, W, ix, iy);
          }
        }
      } else {
        for (int h = 0; h < Hh; ++h) {
          const float* a = p->attn_logits + (((size_t)b * Q + q) * Hh + h) * LP;
          float m = a[0];
          for (int j = 1; j < LP; ++j) m = a[j] > m ? a[j] : m;
          float sm[64], s = 0.f;
          for (int j = 0; j < LP; ++j) { sm[j] = expf(a[j] - m); s += sm[j]; }
          for (int j = 0; j < LP; ++j) sm[j] /= s;
          for (int n = 0; n < N; ++n) {
            const float wc = sigmoidf_(p->cam_logits[(size_t)b * N * Q + (size_t)n * Q + q]);
            for (int pi = 0; pi < P; ++pi) {
              const float* o = p->offsets + ((((size_t)b * Q + q) * Hh + h) * P + pi) * 3;
              float u, v;
              int ok = project(p->lidar2img + ((size_t)b * N + n) * 16, X0 + o[0], Y0 + o[1], Z0 + o[2],
                               p->img_w, p->img_h, 1, &u, &v);
              if (p->mask) p->mask[((((size_t)b * N + n) * Q + q) * Hh + h) * P + pi] = (uint8_t)ok;
              if (!ok) continue;
              float gx = 2.f * u - 1.f, gy = 2.f * v - 1.f;
              for (int l = 0; l < L; ++l) {
                const int H = p->level_h[l], W = p->level_w[l];
                const float wt = sm[l * P + pi] * wc;
                float ix = (gx + 1.f) * (W * 0.5f) - 0.5f, iy = (gy + 1.f) * (H * 0.5f) - 0.5f;
                const float* img = p->value[l] + ((size_t)b * N + n) * C * H * W;
                for (int c = 0; c < Ch; ++c) out[h * Ch + c] += wt * bilinear(img, h * Ch + c, H, W, ix, iy);
              }
            }
          }
        }
      }
    }
  }
  return 0;
}
