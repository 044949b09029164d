// Backward candidate list, per-item backward records and the vector reduction, shared by the
// atomics-based backward (xview_bwd.cu) and the sorted owner-computes backward (xview_bwd_sorted.cu).
#pragma once
#include "xview_common.cuh"

namespace gd4d {

struct __align__(16) CandB {
  float u, v, den, w;   // w = sigmoid(cam logit) (C) or 1 (A)
  int np;               // n<<8 | p
  float du, dv, cg;     // accumulators: dL/du, dL/dv, sum_l sm[l,p]*(s.g)
  __device__ __forceinline__ static CandB make(const Projected& pr, int n, int pi, float wc) {
    CandB c;
    c.u = pr.u; c.v = pr.v; c.den = pr.den; c.w = wc; c.np = (n << 8) | pi;
    c.du = c.dv = c.cg = 0.f;
    return c;
  }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a),
               "f"(b), "f"(c), "f"(d)
               : "memory");
}

// Backward record of one (candidate, level) item, built by ONE lane (see xview_fwd.cu).
// Corner offsets are clamped into the map (always valid addresses); every coefficient
// already carries the in-bounds mask, so out-of-map corners contribute exact zeros:
//   s      = sum_c w_c  f_c          (bilinear sample)
//   W*ds/dix = sum_c ax_c f_c ,  H*ds/diy = sum_c ay_c f_c
struct __align__(16) RecB {
  long long o00, o01, o10, o11;   // element offsets of the 4 corner rows (without lane offset)
  float w00, w01, w10, w11;
  float ax00, ax01, ax10, ax11;
  float ay00, ay01, ay10, ay11;
  float wt, smw, cw;              // total weight, softmax weight, camera weight
  int meta;                       // active<<31 | k<<16 | slot<<8 | l
};

template <int MODE, typename VT, bool WIDE>
__device__ __forceinline__ RecB build_record_bwd(const gd4d_xview_params& p, const CandB* cands,
                                                 const float* sw, int item, int total,
                                                 const WarpCtx& w) {
  RecB r;
  const bool active = item < total;
  const int it = active ? item : 0;
  const int k = it / p.L;
  const int l = it - k * p.L;
  const int np = cands[k].np;
  const float cu = cands[k].u, cv = cands[k].v, cw = cands[k].w;
  const int n = np >> 8;
  const int pi = np & 0xff;
  float wt, smw = 0.f;
  if (MODE == GD4D_MODE_C) {
    smw = sw[l * p.P + pi];
    wt = smw * cw;
  } else {
    const float* a = p.attn_logits + (static_cast<size_t>(w.bq) * p.N + n) * p.P * p.L + l;
    wt = 0.f;
    for (int pp = 0; pp < p.P; ++pp) wt += sigmoidf_(__ldg(a + pp * p.L));
  }
  if (!active) wt = 0.f;
  const int W = p.level_w[l], H = p.level_h[l];
  const float ix = to_pixel(to_grid<MODE>(cu), static_cast<float>(W));
  const float iy = to_pixel(to_grid<MODE>(cv), static_cast<float>(H));
  const Footprint f = footprint(ix, iy, W, H);
  const float m00 = f.in00 ? 1.f : 0.f, m01 = f.in01 ? 1.f : 0.f;
  const float m10 = f.in10 ? 1.f : 0.f, m11 = f.in11 ? 1.f : 0.f;
  const float fW = static_cast<float>(W), fH = static_cast<float>(H);
  r.w00 = m00 * (1.f - f.tx) * (1.f - f.ty); r.w01 = m01 * f.tx * (1.f - f.ty);
  r.w10 = m10 * (1.f - f.tx) * f.ty;         r.w11 = m11 * f.tx * f.ty;
  r.ax00 = -m00 * (1.f - f.ty) * fW; r.ax01 = m01 * (1.f - f.ty) * fW;
  r.ax10 = -m10 * f.ty * fW;         r.ax11 = m11 * f.ty * fW;
  r.ay00 = -m00 * (1.f - f.tx) * fH; r.ay10 = m10 * (1.f - f.tx) * fH;
  r.ay01 = -m01 * f.tx * fH;         r.ay11 = m11 * f.tx * fH;
  r.wt = wt; r.smw = smw; r.cw = cw;
  const int slot = (MODE == GD4D_MODE_C) ? (l * p.P + pi) : 0;  // softmax slot (<= 63); unused in mode A
  r.meta = static_cast<int>((active ? 0x80000000u : 0u) | (static_cast<unsigned>(k) << 16) |
                            (static_cast<unsigned>(slot) << 8) | static_cast<unsigned>(l));
  const int x0 = f.x0, x1 = f.x0 + 1, y0 = f.y0, y1 = f.y0 + 1;
  const long long img = static_cast<long long>(w.b) * p.N + n;
  const long long base = img * H * W * p.C + (WIDE ? 0 : w.h * kHeadDim);
  // out-of-map corners (all coefficients 0, REDs skipped) and padding records gather the zero row
  // (xview_common.cuh) -- its element offset from this level's base; both are 16-byte aligned
  const long long z = (reinterpret_cast<const char*>(g_zero_row) - static_cast<const char*>(p.value[l])) /
                      static_cast<long long>(sizeof(VT));
  r.o00 = (active && f.in00) ? base + (static_cast<long long>(y0) * W + x0) * p.C : z;
  r.o01 = (active && f.in01) ? base + (static_cast<long long>(y0) * W + x1) * p.C : z;
  r.o10 = (active && f.in10) ? base + (static_cast<long long>(y1) * W + x0) * p.C : z;
  r.o11 = (active && f.in11) ? base + (static_cast<long long>(y1) * W + x1) * p.C : z;
  return r;
}

}  // namespace gd4d
