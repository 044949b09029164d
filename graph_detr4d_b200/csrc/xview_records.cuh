// Forward candidate list + per-item gather records, shared by the LDG and the TMA
// (cp.async.bulk) forward kernels.
#pragma once
#include "xview_common.cuh"

namespace gd4d {

// candidate record kept in shared memory: u, v, camera weight, packed (n<<8 | p)
struct __align__(16) Cand {
  float u, v, w;
  int np;
  __device__ __forceinline__ static Cand make(const Projected& pr, int n, int pi, float wc) {
    Cand c;
    c.u = pr.u; c.v = pr.v; c.w = wc; c.np = (n << 8) | pi;
    return c;
  }
};

// gather record of one (candidate, level) item
struct __align__(16) RecF {
  const char* p00;
  const char* p01;
  const char* p10;
  const char* p11;
  float w00, w01, w10, w11;  // mode C: bilinear weight * attention weight; mode A: bilinear weight only
};                           // (the sample passes through nan_to_num before its weight); 0 when out of the map

template <int MODE, typename VT, bool WIDE>
__device__ __forceinline__ RecF build_record(const gd4d_xview_params& p, const Cand* cands,
                                             const float* sw, int item, int total, const WarpCtx& w,
                                             float& wsum_lane, float& wt_out) {
  RecF r;
  const bool active = item < total;
  const int it = active ? item : 0;
  const int k = it / p.L;
  const int l = it - k * p.L;
  const Cand c = cands[k];
  const int n = c.np >> 8;
  const int pi = c.np & 0xff;
  float wt;
  if (MODE == GD4D_MODE_C) {
    wt = sw[l * p.P + pi] * c.w;
  } else {  // mode A: sum_p sigmoid(a[b,q,n,p,l])  (one sample broadcast against P weights)
    const float* a = p.attn_logits + (static_cast<size_t>(w.bq) * p.N + n) * p.P * p.L + l;
    wt = 0.f;
    for (int pp = 0; pp < p.P; ++pp) wt += sigmoidf_(__ldg(a + pp * p.L));
  }
  if (!active) wt = 0.f;
  const int W = p.level_w[l], H = p.level_h[l];
  const float ix = to_pixel(to_grid<MODE>(c.u), static_cast<float>(W));
  const float iy = to_pixel(to_grid<MODE>(c.v), static_cast<float>(H));
  const Footprint f = footprint(ix, iy, W, H);
  const float b00 = (1.f - f.tx) * (1.f - f.ty), b01 = f.tx * (1.f - f.ty);
  const float b10 = (1.f - f.tx) * f.ty, b11 = f.tx * f.ty;
  const float i00 = f.in00 ? b00 : 0.f, i01 = f.in01 ? b01 : 0.f;
  const float i10 = f.in10 ? b10 : 0.f, i11 = f.in11 ? b11 : 0.f;
  wt_out = wt;
  if (MODE == GD4D_MODE_C) {
    r.w00 = wt * i00; r.w01 = wt * i01; r.w10 = wt * i10; r.w11 = wt * i11;
  } else {
    r.w00 = i00; r.w01 = i01; r.w10 = i10; r.w11 = i11;
  }
  wsum_lane = fmaf(wt, (i00 + i01) + (i10 + i11), wsum_lane);
  const int x0 = f.x0, x1 = f.x0 + 1, y0 = f.y0, y1 = f.y0 + 1;
  const size_t img = static_cast<size_t>(w.b) * p.N + n;
  const size_t rowb = static_cast<size_t>(p.C) * sizeof(VT);
  const char* base = static_cast<const char*>(p.value[l]) + img * H * W * rowb +
                     (WIDE ? 0 : static_cast<size_t>(w.h) * kHeadDim * sizeof(VT));
  const size_t r0 = static_cast<size_t>(y0) * W, r1 = static_cast<size_t>(y1) * W;
  // out-of-map corners (weight 0) and padding records gather the zero row: always a valid address,
  // never a pixel the reference would not read
  const char* z = reinterpret_cast<const char*>(g_zero_row);
  r.p00 = (active && f.in00) ? base + (r0 + x0) * rowb : z;
  r.p01 = (active && f.in01) ? base + (r0 + x1) * rowb : z;
  r.p10 = (active && f.in10) ? base + (r1 + x0) * rowb : z;
  r.p11 = (active && f.in11) ? base + (r1 + x1) * rowb : z;
  return r;
}

}  // namespace gd4d
