// Shared device helpers for the cross-view sampling kernels (sm_100a only).
//
// Geometry follows the reference op-for-op so the projection mask is bit-exact
// against the CPU oracle:
//   X = r*span + lo                      detr3d_transformer.py:405-407
//   cam = ((M0*X + M1*Y) + M2*Z) + M3    detr3d_transformer.py:409-414 (sequential, no FMA)
//   valid = cz > 1e-5 ; den = max(cz,1e-5); u = (cx/den)/W_img ; v = (cy/den)/H_img   :415-420
// Every op is an explicit round-to-nearest intrinsic so nvcc cannot contract
// mul+add into FMA or turn the divisions into reciprocal multiplies.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gd4d_xview.h"

namespace gd4d {

constexpr int kWarpsPerCta = 8;
constexpr int kHeadDim = 32;       // channels per head slice (C / Hh)
constexpr int kMaxLP = 64;         // L*P logits per head kept in shared memory
constexpr float kEps = 1e-5f;

struct Projected {
  float u, v;      // normalised image coords (divided by the unpadded image size)
  float cx, cy;    // camera-plane numerators   (backward only)
  float den;       // max(cz, eps)              (backward only)
  bool depth_ok;   // cz > eps
};

__device__ __forceinline__ Projected project_point(const float* __restrict__ M, float X, float Y,
                                                   float Z, float img_w, float img_h) {
  Projected r;
  const float m0 = __ldg(M + 0), m1 = __ldg(M + 1), m2 = __ldg(M + 2), m3 = __ldg(M + 3);
  const float m4 = __ldg(M + 4), m5 = __ldg(M + 5), m6 = __ldg(M + 6), m7 = __ldg(M + 7);
  const float m8 = __ldg(M + 8), m9 = __ldg(M + 9), m10 = __ldg(M + 10), m11 = __ldg(M + 11);
  r.cx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, X), __fmul_rn(m1, Y)), __fmul_rn(m2, Z)), m3);
  r.cy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4, X), __fmul_rn(m5, Y)), __fmul_rn(m6, Z)), m7);
  const float cz =
      __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m8, X), __fmul_rn(m9, Y)), __fmul_rn(m10, Z)), m11);
  r.depth_ok = cz > kEps;
  r.den = fmaxf(cz, kEps);
  r.u = __fdiv_rn(__fdiv_rn(r.cx, r.den), img_w);
  r.v = __fdiv_rn(__fdiv_rn(r.cy, r.den), img_h);
  return r;
}

// In-image test.  Mode A / V2 test the grid coordinate g=(u-0.5)*2 against (-1,1)
// (detr3d_transformer.py:421-425); mode C tests u,v against (0,1)
// (deform3d_cross_attn.py:249-252).
template <int MODE>
__device__ __forceinline__ bool in_image(float u, float v) {
  if (MODE == GD4D_MODE_C) {
    return (u > 0.f) & (u < 1.f) & (v > 0.f) & (v < 1.f);
  } else {
    const float gx = __fmul_rn(__fsub_rn(u, 0.5f), 2.f);
    const float gy = __fmul_rn(__fsub_rn(v, 0.5f), 2.f);
    return (gx > -1.f) & (gx < 1.f) & (gy > -1.f) & (gy < 1.f);
  }
}

// Normalised coord -> grid coord in [-1,1] as the oracle forms it:
//   A/V2: g = (u - 0.5) * 2            detr3d_transformer.py:421
//   C   : g = 2*u - 1                  mmcv multi_scale_deformable_attn_pytorch
template <int MODE>
__device__ __forceinline__ float to_grid(float u) {
  if (MODE == GD4D_MODE_C) return __fsub_rn(__fmul_rn(2.f, u), 1.f);
  return __fmul_rn(__fsub_rn(u, 0.5f), 2.f);
}

// grid_sample un-normalisation, align_corners=False: ix = (g+1)*(size/2) - 0.5
__device__ __forceinline__ float to_pixel(float g, float size) {
  return __fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), size * 0.5f), 0.5f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- 16-byte channel-slice loads ------------------------------------------------
template <typename VT>
struct Slice;  // VEC = channels per 16-byte lane load

template <>
struct Slice<float> {
  static constexpr int VEC = 4;
  static constexpr int LANES = kHeadDim / VEC;  // 8 lanes cover one 128-byte head slice
  __device__ __forceinline__ static void load(const float* p, bool pred, float (&v)[VEC]) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pred) t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};

template <>
struct Slice<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static constexpr int LANES = kHeadDim / VEC;  // 4 lanes cover one 64-byte head slice
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, bool pred, float (&v)[VEC]) {
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (pred) t = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
    v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
  }
};

// Bilinear footprint of one sample in one level (zeros padding, not clamped).
struct Footprint {
  int x0, y0;
  float tx, ty;           // ix - x0, iy - y0
  bool in00, in01, in10, in11;  // (y,x): 00=(y0,x0) 01=(y0,x1) 10=(y1,x0) 11=(y1,x1)
};

__device__ __forceinline__ Footprint footprint(float ix, float iy, int W, int H) {
  Footprint f;
  const float fx = floorf(ix), fy = floorf(iy);
  f.tx = ix - fx;
  f.ty = iy - fy;
  // clamp before the int conversion so absurd coordinates cannot overflow
  f.x0 = static_cast<int>(fminf(fmaxf(fx, -2.f), static_cast<float>(W)));
  f.y0 = static_cast<int>(fminf(fmaxf(fy, -2.f), static_cast<float>(H)));
  const bool xin0 = (f.x0 >= 0) & (f.x0 < W), xin1 = (f.x0 + 1 >= 0) & (f.x0 + 1 < W);
  const bool yin0 = (f.y0 >= 0) & (f.y0 < H), yin1 = (f.y0 + 1 >= 0) & (f.y0 + 1 < H);
  f.in00 = yin0 & xin0; f.in01 = yin0 & xin1; f.in10 = yin1 & xin0; f.in11 = yin1 & xin1;
  return f;
}

struct LaunchGeom {
  int grid, block, smem, cand_cap;
};

}  // namespace gd4d
