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

// 2 KB of zeros.  Out-of-map bilinear corners and padding records of the branch-free gathers point
// here instead of at a clamped in-map pixel: the reference's zeros padding (F.grid_sample, mmcv MSDA)
// never READS a pixel outside the map, and weight 0 times a non-finite border pixel would be NaN.
static __device__ __align__(16) unsigned char g_zero_row[2048];

// torch.nan_to_num (detr3d_transformer.py:378, :619): NaN -> 0, +-Inf -> +-FLT_MAX
__device__ __forceinline__ float nan_to_num_(float x) {
  if (x != x) return 0.f;
  if (fabsf(x) == INFINITY) return copysignf(3.402823466e+38f, x);
  return x;
}

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
// Loads are volatile inline PTX so that nvcc keeps them in program order, and
// `pin()` is an empty volatile asm that "touches" the loaded registers: arithmetic
// that consumes them cannot be hoisted above it.  Together they make the compiler
// issue a whole batch of independent gathers BEFORE the first FMA waits on one --
// left alone it interleaves load/consume to save registers and the warp has only
// 2-4 gathers in flight (seen in the r1 SASS), which is what bounds this kernel.
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p, bool pred) {
  uint4 t;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\t"
      "mov.b32 %1, 0;\n\t"
      "mov.b32 %2, 0;\n\t"
      "mov.b32 %3, 0;\n\t"
      "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t"
      "}"
      : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
      : "l"(p), "r"(static_cast<int>(pred)));
  return t;
}

// unpredicated variant: the caller guarantees a valid address (clamped corners + zero weight)
__device__ __forceinline__ uint4 ldg_nc_v4_all(const char* p) {
  uint4 t;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
               : "l"(p));
  return t;
}

// L2 prefetch of a row the warp will gather in the NEXT batch: costs no register, so it raises the
// memory-level parallelism beyond what the register file allows (GD4D_FLAG_L2_PREFETCH).
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ void pin(uint4& a, uint4& b, uint4& c, uint4& d) {
  asm volatile("" : "+r"(a.x), "+r"(a.y), "+r"(a.z), "+r"(a.w), "+r"(b.x), "+r"(b.y), "+r"(b.z),
                    "+r"(b.w), "+r"(c.x), "+r"(c.y), "+r"(c.z), "+r"(c.w), "+r"(d.x), "+r"(d.y),
                    "+r"(d.z), "+r"(d.w));
}

template <typename VT>
struct Slice;  // VEC = channels per 16-byte lane load

template <>
struct Slice<float> {
  static constexpr int VEC = 4;
  __device__ __forceinline__ static void unpack(const uint4& t, float* v) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y);
    v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
};

template <>
struct Slice<__nv_bfloat16> {
  static constexpr int VEC = 8;
  __device__ __forceinline__ static void unpack(const uint4& t, float* v) {
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
  int nv;  // wide mode: 16-byte vectors per lane per corner (C*elem/512), else 1
};

// Per-warp work coordinates
struct WarpCtx {
  int b, q, h, bq, lane;
  float X0, Y0, Z0;  // reference point in metres
};

// Work distribution.  Static: warp gw = blockIdx*8 + warp handles item gw (one pass).
// Dynamic (p.sched != NULL): persistent grid, every warp claims the next (b,q,head) item
// from a global counter until none are left; the last warp to leave resets the counter.
struct WorkIter {
  long long total;
  long long next;     // static mode: this warp's single item (or -1 when consumed)
  bool dynamic;
};

__device__ __forceinline__ void work_begin(const gd4d_xview_params& p, WorkIter& it) {
  it.total = static_cast<long long>(p.B) * p.Q * p.Hh;
  it.dynamic = p.sched != nullptr;
  it.next = static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
}

__device__ __forceinline__ bool work_next(const gd4d_xview_params& p, WorkIter& it, WarpCtx& w) {
  long long gw;
  w.lane = threadIdx.x & 31;
  if (it.dynamic) {
    unsigned v = 0;
    if (w.lane == 0) v = atomicAdd(p.sched, 1u);
    gw = __shfl_sync(0xffffffffu, v, 0);
  } else {
    gw = it.next;
    it.next = it.total;  // one item per warp
  }
  if (gw >= it.total) return false;
  w.h = static_cast<int>(gw % p.Hh);
  w.bq = static_cast<int>(gw / p.Hh);
  w.b = w.bq / p.Q;
  w.q = w.bq - w.b * p.Q;
  const float* rp = p.ref + static_cast<size_t>(w.bq) * 3;
  w.X0 = __fadd_rn(__fmul_rn(__ldg(rp + 0), p.pc_span[0]), p.pc_lo[0]);
  w.Y0 = __fadd_rn(__fmul_rn(__ldg(rp + 1), p.pc_span[1]), p.pc_lo[1]);
  w.Z0 = __fadd_rn(__fmul_rn(__ldg(rp + 2), p.pc_span[2]), p.pc_lo[2]);
  return true;
}

__device__ __forceinline__ void work_end(const gd4d_xview_params& p, const WorkIter& it,
                                         int warps_per_cta = kWarpsPerCta) {
  if (!it.dynamic) return;
  if ((threadIdx.x & 31) == 0) {
    const unsigned warps = gridDim.x * warps_per_cta;
    const unsigned done = atomicAdd(p.sched + 1, 1u);
    if (done == warps - 1) {  // every warp has made its final (failing) claim: safe to reset
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

// Mode C generator outputs (attn_logits / offsets / cam_logits and their gradients): dense
// per-tensor layouts when gen_stride == 0, else column blocks of one row-major
// (B*Q, gen_stride) matrix (include/gd4d_xview.h).
__device__ __forceinline__ size_t attn_row_off(const gd4d_xview_params& p, const WarpCtx& w) {
  const int LP = p.L * p.P;
  return p.gen_stride > 0 ? static_cast<size_t>(w.bq) * p.gen_stride + w.h * LP
                          : (static_cast<size_t>(w.bq) * p.Hh + w.h) * LP;
}
__device__ __forceinline__ size_t offsets_row_off(const gd4d_xview_params& p, const WarpCtx& w) {
  return p.gen_stride > 0 ? static_cast<size_t>(w.bq) * p.gen_stride + w.h * p.P * 3
                          : (static_cast<size_t>(w.bq) * p.Hh + w.h) * p.P * 3;
}
// the reference views the (B,Q,N) Linear output as (B,N,Q): weight(n,q) = flat[n*Q + q]
__device__ __forceinline__ size_t cam_off(const gd4d_xview_params& p, const WarpCtx& w, int n) {
  const size_t e = static_cast<size_t>(n) * p.Q + w.q;
  if (p.gen_stride > 0)
    return (static_cast<size_t>(w.b) * p.Q + e / p.N) * p.gen_stride + e % p.N;
  return static_cast<size_t>(w.b) * p.N * p.Q + e;
}

// softmax over the head's L*P (<= 64) logits into sw[0..64) (zeros past L*P)
__device__ __forceinline__ void head_softmax(const gd4d_xview_params& p, const WarpCtx& w, float* sw) {
  const int LP = p.L * p.P;
  const int lane = w.lane;
  const float* a = p.attn_logits + attn_row_off(p, w);
  const float x0 = lane < LP ? __ldg(a + lane) : -INFINITY;
  const float x1 = lane + 32 < LP ? __ldg(a + lane + 32) : -INFINITY;
  float m = fmaxf(x0, x1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e0 = lane < LP ? expf(x0 - m) : 0.f;
  const float e1 = lane + 32 < LP ? expf(x1 - m) : 0.f;
  float s = e0 + e1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  sw[lane] = e0 / s;
  sw[lane + 32] = e1 / s;
}

// Phase 1 of both kernels: the 32 lanes project the warp's N*P candidate points in
// parallel and ballot-compact the valid ones into `cands` (shared memory).
// CandT must provide:  static CandT make(const Projected&, int n, int pi, float wc).
template <int MODE, typename CandT>
__device__ __forceinline__ int build_candidates(const gd4d_xview_params& p, const WarpCtx& w,
                                                CandT* cands, bool write_mask) {
  const int PP = (MODE == GD4D_MODE_C) ? p.P : 1;  // mode A: one centre point per camera
  const int ncand = p.N * PP;
  int nvalid = 0;
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + w.lane;
    bool valid = false;
    Projected pr;
    pr.u = pr.v = pr.cx = pr.cy = 0.f; pr.den = 1.f; pr.depth_ok = false;
    float wc = 1.f;
    int n = 0, pi = 0;
    if (c < ncand) {
      n = c / PP;
      pi = c - n * PP;
      float X = w.X0, Y = w.Y0, Z = w.Z0;
      if (MODE == GD4D_MODE_C) {
        const float* o = p.offsets + offsets_row_off(p, w) + pi * 3;
        X = __fadd_rn(X, __ldg(o + 0));
        Y = __fadd_rn(Y, __ldg(o + 1));
        Z = __fadd_rn(Z, __ldg(o + 2));
      }
      const float* M = p.lidar2img + (static_cast<size_t>(w.b) * p.N + n) * 16;
      pr = project_point(M, X, Y, Z, p.img_w, p.img_h);
      valid = pr.depth_ok & in_image<MODE>(pr.u, pr.v);
      if (write_mask) {
        if (MODE == GD4D_MODE_C)
          p.mask[(((static_cast<size_t>(w.b) * p.N + n) * p.Q + w.q) * p.Hh + w.h) * p.P + pi] = valid;
        else if (w.h == 0)
          p.mask[static_cast<size_t>(w.bq) * p.N + n] = valid;
      }
      if (valid && MODE == GD4D_MODE_C)  // reference views (B,Q,N) memory as (B,N,Q): flat[n*Q+q]
        wc = sigmoidf_(__ldg(p.cam_logits + cam_off(p, w, n)));
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (valid) cands[nvalid + __popc(bal & ((1u << w.lane) - 1u))] = CandT::make(pr, n, pi, wc);
    nvalid += __popc(bal);
  }
  __syncwarp();
  return nvalid;
}

}  // namespace gd4d
