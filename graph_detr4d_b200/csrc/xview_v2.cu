// Mode V2: Detr3DCrossAttenV2 (detr3d_transformer.py:441-709) -- registered in the
// reference, used by no config; built for coverage of SURVEY.md 8a row a6.
//
// Semantics reproduced (incl. quirks, each checked against the executed reference):
//   * one CENTRE projection per (query, camera), mask on g=(u-0.5)*2 in (-1,1)   :667-684
//   * per (camera, head, level, point) 2D offsets added in GRID units:
//       loc = g + offset / [W_l, H_l]   (so one offset unit = half a pixel)     :698-700
//   * softmax over the L*P logits of each (query, camera, head)                  :602-605
//   * the weight tensor is (..,N,L,P) but the sample tensor is (..,N,P,L) (:611 vs :709):
//     sample (point p, level l) is multiplied by softmax_flat[p*P + l]; the reference only
//     runs when L == P, and so does this kernel (UNSUPPORTED otherwise)
//   * head h samples its own 32-channel slice of the RAW maps (no value_proj)    :692
// One warp per (b, q, head); cameras with a valid centre are walked one at a time (1.08 on
// average), their L*P items spread over the lane groups exactly as in xview_fwd.cu (narrow).
#include "xview_common.cuh"

namespace gd4d {

struct __align__(16) CandV {
  float u, v, den, w;
  int np;
  float du, dv, pad;
  __device__ __forceinline__ static CandV make(const Projected& pr, int n, int pi, float wc) {
    CandV c;
    c.u = pr.u; c.v = pr.v; c.den = pr.den; c.w = wc; c.np = (n << 8) | pi;
    c.du = c.dv = c.pad = 0.f;
    return c;
  }
};

__device__ __forceinline__ void red_add_v4_(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a),
               "f"(b), "f"(c), "f"(d)
               : "memory");
}

// softmax of the LP (<=64) logits at `a` into sw[0..64)
__device__ __forceinline__ void softmax64(const float* a, int LP, int lane, float* sw) {
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

struct V2Item {
  int l, pi, widx;      // level, point, index into the softmax vector (p*P + l: see header)
  float ix, iy;
};

__device__ __forceinline__ V2Item v2_item(const gd4d_xview_params& p, const float* off, int j,
                                          float gx, float gy) {
  V2Item it;
  it.l = j / p.P;
  it.pi = j - it.l * p.P;
  it.widx = it.pi * p.P + it.l;
  const float ox = __ldg(off + 2 * j), oy = __ldg(off + 2 * j + 1);
  const float W = static_cast<float>(p.level_w[it.l]), H = static_cast<float>(p.level_h[it.l]);
  it.ix = to_pixel(__fadd_rn(gx, __fdiv_rn(ox, W)), W);
  it.iy = to_pixel(__fadd_rn(gy, __fdiv_rn(oy, H)), H);
  return it;
}

template <typename VT, int LANES, bool BWD>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
xview_v2_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int GROUPS = 32 / LANES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = lane / LANES;
  const int sub = lane % LANES;
  const size_t warp_bytes = sizeof(float) * kMaxLP * 5 + sizeof(CandV) * cand_cap;
  float* sw = reinterpret_cast<float*>(smem_raw + warp * warp_bytes);
  float* gsum = sw + kMaxLP;
  CandV* cands = reinterpret_cast<CandV*>(gsum + 4 * kMaxLP);
  const int LP = p.L * p.P;

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    const int nvalid = build_candidates<GD4D_MODE_V2, CandV>(p, w, cands, !BWD && p.mask != nullptr);
    float acc[VEC];
    float g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    if (BWD) {
      const float* go = p.grad_out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim + sub * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(go + i));
        g[i] = t.x; g[i + 1] = t.y; g[i + 2] = t.z; g[i + 3] = t.w;
      }
    }
    float rX = 0.f, rY = 0.f, rZ = 0.f;
    for (int k = 0; k < nvalid; ++k) {
      const int n = cands[k].np >> 8;
      const float gx = to_grid<GD4D_MODE_V2>(cands[k].u), gy = to_grid<GD4D_MODE_V2>(cands[k].v);
      const size_t qnh = (static_cast<size_t>(w.bq) * p.N + n) * p.Hh + w.h;
      softmax64(p.attn_logits + qnh * LP, LP, lane, sw);
      if (BWD) { gsum[lane] = 0.f; gsum[lane + 32] = 0.f; }
      __syncwarp();
      const float* off = p.offsets + qnh * LP * 2;
      const size_t img = static_cast<size_t>(w.b) * p.N + n;
      float du = 0.f, dv = 0.f;
      for (int j0 = 0; j0 < LP; j0 += GROUPS) {
        const int j = j0 + grp;
        const bool active = j < LP;
        const V2Item it = v2_item(p, off, active ? j : 0, gx, gy);
        const int W = p.level_w[it.l], H = p.level_h[it.l];
        const Footprint f = footprint(it.ix, it.iy, W, H);
        const float wt = active ? sw[it.widx] : 0.f;
        const VT* base = static_cast<const VT*>(p.value[it.l]);
        const size_t e00 = ((img * H + f.y0) * W + f.x0) * p.C + static_cast<size_t>(w.h) * kHeadDim + sub * VEC;
        const size_t rowst = static_cast<size_t>(W) * p.C;
        const bool a00 = active & f.in00, a01 = active & f.in01, a10 = active & f.in10, a11 = active & f.in11;
        uint4 r00 = ldg_nc_v4(base + e00, a00), r01 = ldg_nc_v4(base + e00 + p.C, a01);
        uint4 r10 = ldg_nc_v4(base + e00 + rowst, a10), r11 = ldg_nc_v4(base + e00 + rowst + p.C, a11);
        pin(r00, r01, r10, r11);
        float c00[VEC], c01[VEC], c10[VEC], c11[VEC];
        Slice<VT>::unpack(r00, c00); Slice<VT>::unpack(r01, c01);
        Slice<VT>::unpack(r10, c10); Slice<VT>::unpack(r11, c11);
        const float w00 = (1.f - f.tx) * (1.f - f.ty), w01 = f.tx * (1.f - f.ty);
        const float w10 = (1.f - f.tx) * f.ty, w11 = f.tx * f.ty;
        if (!BWD) {
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            acc[i] = fmaf(wt, nan_to_num_(w00 * c00[i] + w01 * c01[i] + w10 * c10[i] + w11 * c11[i]), acc[i]);  // :619
        } else {
          float* gv = p.grad_value[it.l];
          if (gv != nullptr && wt != 0.f) {
#pragma unroll
            for (int i = 0; i < VEC; i += 4) {
              if (a00) red_add_v4_(gv + e00 + i, wt * w00 * g[i], wt * w00 * g[i + 1], wt * w00 * g[i + 2], wt * w00 * g[i + 3]);
              if (a01) red_add_v4_(gv + e00 + p.C + i, wt * w01 * g[i], wt * w01 * g[i + 1], wt * w01 * g[i + 2], wt * w01 * g[i + 3]);
              if (a10) red_add_v4_(gv + e00 + rowst + i, wt * w10 * g[i], wt * w10 * g[i + 1], wt * w10 * g[i + 2], wt * w10 * g[i + 3]);
              if (a11) red_add_v4_(gv + e00 + rowst + p.C + i, wt * w11 * g[i], wt * w11 * g[i + 1], wt * w11 * g[i + 2], wt * w11 * g[i + 3]);
            }
          }
          float sdot = 0.f, dxdot = 0.f, dydot = 0.f;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            sdot += g[i] * (w00 * c00[i] + w01 * c01[i] + w10 * c10[i] + w11 * c11[i]);
            dxdot += g[i] * ((c01[i] - c00[i]) * (1.f - f.ty) + (c11[i] - c10[i]) * f.ty);
            dydot += g[i] * ((c10[i] - c00[i]) * (1.f - f.tx) + (c11[i] - c01[i]) * f.tx);
          }
#pragma unroll
          for (int o = 1; o < LANES; o <<= 1) {
            sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
            dxdot += __shfl_xor_sync(0xffffffffu, dxdot, o);
            dydot += __shfl_xor_sync(0xffffffffu, dydot, o);
          }
          if (active && sub == 0) {
            gsum[it.widx] = sdot;                                  // each softmax slot has one item
            du += wt * static_cast<float>(W) * dxdot;              // d ix / d u = W
            dv += wt * static_cast<float>(H) * dydot;
            if (p.grad_offsets != nullptr) {                       // d ix / d off_x = (1/W)*(W/2)
              float* go2 = p.grad_offsets + (qnh * LP + j) * 2;
              atomicAdd(go2, 0.5f * wt * dxdot);
              atomicAdd(go2 + 1, 0.5f * wt * dydot);
            }
          }
        }
      }
      if (BWD) {
        __syncwarp();
        if (p.grad_attn_logits != nullptr) {
          const float s0 = sw[lane], s1 = sw[lane + 32];
          const float g0 = gsum[lane], g1 = gsum[lane + 32];
          float dot = s0 * g0 + s1 * g1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          float* ga = p.grad_attn_logits + qnh * LP;
          if (lane < LP) atomicAdd(ga + lane, s0 * (g0 - dot));
          if (lane + 32 < LP) atomicAdd(ga + lane + 32, s1 * (g1 - dot));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // du/dv live in the sub==0 lanes of each group
          du += __shfl_xor_sync(0xffffffffu, du, o);
          dv += __shfl_xor_sync(0xffffffffu, dv, o);
        }
        const CandV cd = cands[k];
        const float* M = p.lidar2img + img * 16;
        const float dcx = du / (cd.den * p.img_w), dcy = dv / (cd.den * p.img_h);
        const float dcz = -(du * cd.u + dv * cd.v) / cd.den;
        rX += __ldg(M + 0) * dcx + __ldg(M + 4) * dcy + __ldg(M + 8) * dcz;
        rY += __ldg(M + 1) * dcx + __ldg(M + 5) * dcy + __ldg(M + 9) * dcz;
        rZ += __ldg(M + 2) * dcx + __ldg(M + 6) * dcy + __ldg(M + 10) * dcz;
      }
      __syncwarp();
    }
    if (!BWD) {
#pragma unroll
      for (int o = LANES; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
      }
      if (grp == 0) {
        float* o = p.out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim + sub * VEC;
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
    } else if (p.grad_ref != nullptr && lane == 0 && nvalid > 0) {
      float* gr = p.grad_ref + static_cast<size_t>(w.bq) * 3;
      atomicAdd(gr + 0, rX * p.pc_span[0]);
      atomicAdd(gr + 1, rY * p.pc_span[1]);
      atomicAdd(gr + 2, rZ * p.pc_span[2]);
    }
    __syncwarp();
  }
  work_end(p, wi);
}

template <typename VT, int LANES, bool BWD>
static int launch_v2(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_v2_kernel<VT, LANES, BWD>;
  if (g.smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem) != cudaSuccess)
    return GD4D_ERR_CUDA;
  kern<<<g.grid, g.block, g.smem, stream>>>(p, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int dispatch_v2(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream, bool backward) {
  const bool bf16 = p.value_dtype == GD4D_BF16;
  if (backward)
    return bf16 ? launch_v2<__nv_bfloat16, 4, true>(p, g, stream) : launch_v2<float, 8, true>(p, g, stream);
  return bf16 ? launch_v2<__nv_bfloat16, 4, false>(p, g, stream) : launch_v2<float, 8, false>(p, g, stream);
}

}  // namespace gd4d
