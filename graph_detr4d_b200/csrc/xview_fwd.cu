// Fused project -> mask -> bilinear-sample -> weight -> reduce forward kernel.
//
// One warp per (batch b, query q, head h).  Phase 1: the 32 lanes project the
// warp's N*P candidate points (camera n, graph-offset point p) in parallel and
// ballot-compact the valid ones into a per-warp shared-memory list.  Phase 2:
// the warp walks (valid candidate, level) items; a lane GROUP (8 lanes for fp32,
// 4 for bf16: one 16-byte load per lane covers the head's 32-channel slice)
// owns one item, so every warp-wide load instruction fetches 4 (fp32) or 8
// (bf16) full corner slices, fully coalesced, and two items per group are kept
// in flight (8 independent 16-byte loads per lane).  fp32 accumulation in
// registers, a shuffle reduce across groups, one 128-byte store per (q, head).
//
// Replaces (reference, projects/mmdet3d_plugin/models/utils/):
//   mode A  detr3d_transformer.py:376-383 + feature_sampling :397-438
//   mode C  deform3d_cross_attn.py:227-258, 274, 281-284, 301-304, 320-324
#include "xview_common.cuh"

namespace gd4d {

// candidate record kept in shared memory: u, v, camera weight, packed (n<<8 | p)
struct __align__(16) Cand {
  float u, v, w;
  int np;
};

template <int MODE, typename VT>
struct ItemLoad {
  static constexpr int VEC = Slice<VT>::VEC;
  float c00[VEC], c01[VEC], c10[VEC], c11[VEC];
  float w00, w01, w10, w11, wt;
};

template <int MODE, typename VT>
__device__ __forceinline__ void item_issue(const gd4d_xview_params& p, const Cand* cands,
                                           const float* sw, int item, int total, int b, int q,
                                           int h, int sub, ItemLoad<MODE, VT>& ld) {
  constexpr int VEC = Slice<VT>::VEC;
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
    const float* a = p.attn_logits + ((static_cast<size_t>(b) * p.Q + q) * p.N + n) * p.P * p.L + l;
    wt = 0.f;
    for (int pp = 0; pp < p.P; ++pp) wt += sigmoidf_(__ldg(a + pp * p.L));
  }
  const int W = p.level_w[l], H = p.level_h[l];
  const float ix = to_pixel(to_grid<MODE>(c.u), static_cast<float>(W));
  const float iy = to_pixel(to_grid<MODE>(c.v), static_cast<float>(H));
  const Footprint f = footprint(ix, iy, W, H);
  ld.w00 = (1.f - f.tx) * (1.f - f.ty);
  ld.w01 = f.tx * (1.f - f.ty);
  ld.w10 = (1.f - f.tx) * f.ty;
  ld.w11 = f.tx * f.ty;
  ld.wt = active ? wt : 0.f;
  const VT* base = static_cast<const VT*>(p.value[l]);
  const size_t img = static_cast<size_t>(b) * p.N + n;
  const size_t row0 = (img * H + f.y0) * W;
  const size_t choff = static_cast<size_t>(h) * kHeadDim + sub * VEC;
  const VT* p00 = base + (row0 + f.x0) * p.C + choff;
  const VT* p10 = p00 + static_cast<size_t>(W) * p.C;
  Slice<VT>::load(p00, active & f.in00, ld.c00);
  Slice<VT>::load(p00 + p.C, active & f.in01, ld.c01);
  Slice<VT>::load(p10, active & f.in10, ld.c10);
  Slice<VT>::load(p10 + p.C, active & f.in11, ld.c11);
}

template <int MODE, typename VT>
__device__ __forceinline__ void item_consume(const ItemLoad<MODE, VT>& ld,
                                             float (&acc)[Slice<VT>::VEC]) {
#pragma unroll
  for (int i = 0; i < Slice<VT>::VEC; ++i) {
    const float s = ld.w00 * ld.c00[i] + ld.w01 * ld.c01[i] + ld.w10 * ld.c10[i] + ld.w11 * ld.c11[i];
    acc[i] = fmaf(ld.wt, s, acc[i]);
  }
}

template <int MODE, typename VT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
xview_fwd_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int LANES = Slice<VT>::LANES;  // lanes per head slice
  constexpr int GROUPS = 32 / LANES;       // items per warp-wide load

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = lane / LANES;
  const int sub = lane % LANES;
  const size_t warp_bytes = sizeof(float) * kMaxLP + sizeof(Cand) * cand_cap;
  float* sw = reinterpret_cast<float*>(smem_raw + warp * warp_bytes);
  Cand* cands = reinterpret_cast<Cand*>(sw + kMaxLP);

  const long long gw = static_cast<long long>(blockIdx.x) * kWarpsPerCta + warp;
  const long long total_warps = static_cast<long long>(p.B) * p.Q * p.Hh;
  if (gw >= total_warps) return;
  const int h = static_cast<int>(gw % p.Hh);
  const int bq = static_cast<int>(gw / p.Hh);
  const int b = bq / p.Q;
  const int q = bq - b * p.Q;

  // ---- reference point in metres ------------------------------------------------
  const float* rp = p.ref + static_cast<size_t>(bq) * 3;
  const float X0 = __fadd_rn(__fmul_rn(__ldg(rp + 0), p.pc_span[0]), p.pc_lo[0]);
  const float Y0 = __fadd_rn(__fmul_rn(__ldg(rp + 1), p.pc_span[1]), p.pc_lo[1]);
  const float Z0 = __fadd_rn(__fmul_rn(__ldg(rp + 2), p.pc_span[2]), p.pc_lo[2]);

  // ---- softmax over the head's L*P logits (mode C) ---------------------------------
  if (MODE == GD4D_MODE_C) {
    const int LP = p.L * p.P;
    const float* a = p.attn_logits + (static_cast<size_t>(bq) * p.Hh + h) * LP;
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
    if (lane < LP) sw[lane] = e0 / s;
    if (lane + 32 < LP) sw[lane + 32] = e1 / s;
  }

  // ---- phase 1: project candidates, compact the valid ones ----------------------------
  const int PP = (MODE == GD4D_MODE_C) ? p.P : 1;  // mode A: one centre point per camera
  const int ncand = p.N * PP;
  int nvalid = 0;
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    bool valid = false;
    float u = 0.f, v = 0.f, wc = 1.f;
    int n = 0, pi = 0;
    if (c < ncand) {
      n = c / PP;
      pi = c - n * PP;
      float X = X0, Y = Y0, Z = Z0;
      if (MODE == GD4D_MODE_C) {
        const float* o = p.offsets + ((static_cast<size_t>(bq) * p.Hh + h) * p.P + pi) * 3;
        X = __fadd_rn(X, __ldg(o + 0));
        Y = __fadd_rn(Y, __ldg(o + 1));
        Z = __fadd_rn(Z, __ldg(o + 2));
      }
      const float* M = p.lidar2img + (static_cast<size_t>(b) * p.N + n) * 16;
      const Projected pr = project_point(M, X, Y, Z, p.img_w, p.img_h);
      u = pr.u;
      v = pr.v;
      valid = pr.depth_ok & in_image<MODE>(u, v);
      if (p.mask != nullptr) {
        if (MODE == GD4D_MODE_C)
          p.mask[(((static_cast<size_t>(b) * p.N + n) * p.Q + q) * p.Hh + h) * p.P + pi] = valid;
        else if (h == 0)
          p.mask[static_cast<size_t>(bq) * p.N + n] = valid;
      }
      if (valid && MODE == GD4D_MODE_C)  // reference views (B,Q,N) memory as (B,N,Q): flat[n*Q+q]
        wc = sigmoidf_(__ldg(p.cam_logits + static_cast<size_t>(b) * p.N * p.Q +
                             static_cast<size_t>(n) * p.Q + q));
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const int pos = nvalid + __popc(bal & ((1u << lane) - 1u));
      Cand cd;
      cd.u = u; cd.v = v; cd.w = wc; cd.np = (n << 8) | pi;
      cands[pos] = cd;
    }
    nvalid += __popc(bal);
  }
  __syncwarp();

  // ---- phase 2: gather + accumulate -------------------------------------------------
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  float bsum = 0.f;  // sum of wt * (in-bounds corner weights): multiplies value_bias
  const int total = nvalid * p.L;
  for (int it0 = 0; it0 < total; it0 += 2 * GROUPS) {
    ItemLoad<MODE, VT> la, lb;
    item_issue<MODE, VT>(p, cands, sw, it0 + grp, total, b, q, h, sub, la);
    item_issue<MODE, VT>(p, cands, sw, it0 + GROUPS + grp, total, b, q, h, sub, lb);
    item_consume<MODE, VT>(la, acc);
    item_consume<MODE, VT>(lb, acc);
  }
  (void)bsum;

  // ---- reduce across lane groups, store the head slice ---------------------------------
#pragma unroll
  for (int o = LANES; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (grp == 0) {
    float* o = p.out + static_cast<size_t>(bq) * p.C + h * kHeadDim + sub * VEC;
#pragma unroll
    for (int i = 0; i < VEC; i += 4)
      *reinterpret_cast<float4*>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  }
}

template <int MODE, typename VT>
static int launch_fwd(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_fwd_kernel<MODE, VT>;
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem);
    if (e != cudaSuccess) return GD4D_ERR_CUDA;
  }
  kern<<<g.grid, g.block, g.smem, stream>>>(p, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int dispatch_forward(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  const bool bf16 = p.value_dtype == GD4D_BF16;
  switch (p.mode) {
    case GD4D_MODE_A:
      return bf16 ? launch_fwd<GD4D_MODE_A, __nv_bfloat16>(p, g, stream)
                  : launch_fwd<GD4D_MODE_A, float>(p, g, stream);
    case GD4D_MODE_C:
      return bf16 ? launch_fwd<GD4D_MODE_C, __nv_bfloat16>(p, g, stream)
                  : launch_fwd<GD4D_MODE_C, float>(p, g, stream);
    default:
      return GD4D_ERR_UNSUPPORTED;
  }
}

}  // namespace gd4d
