// Fused project -> mask -> bilinear-sample -> weight -> reduce forward kernel.
//
// One warp per (batch b, query q, head h).  Phase 1: the 32 lanes project the
// warp's N*P candidate points (camera n, graph-offset point p) in parallel and
// ballot-compact the valid ones into a per-warp shared-memory list.  Phase 2:
// the warp walks (valid candidate, level) items.  A lane GROUP owns one item and
// covers its channel slice with 16-byte loads:
//   narrow (mmcv layout, value already projected, 32-ch head slice):
//       fp32 8 lanes x 1 vector, bf16 4 lanes x 1 vector -> 4 / 8 items per warp-wide load
//   wide   (gather-then-project, all C channels of the raw maps per head):
//       32 lanes x NV vectors (fp32 C=256: NV=2, bf16 C=256: NV=1) -> 1 item per NV loads
// so every warp-wide load instruction fetches whole 64..512-byte runs, fully
// coalesced, and two items per group are kept in flight.  fp32 accumulation in
// registers, a shuffle reduce across groups, coalesced 16-byte stores.
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
  __device__ __forceinline__ static Cand make(const Projected& pr, int n, int pi, float wc) {
    Cand c;
    c.u = pr.u; c.v = pr.v; c.w = wc; c.np = (n << 8) | pi;
    return c;
  }
};

template <typename VT, int NV>
struct ItemLoad {
  static constexpr int PL = Slice<VT>::VEC * NV;  // channels per lane
  uint4 r00[NV], r01[NV], r10[NV], r11[NV];       // raw 16-byte corner runs
  float w00, w01, w10, w11, wt, inb;
};

template <int MODE, typename VT, int LANES, int NV>
__device__ __forceinline__ void item_issue(const gd4d_xview_params& p, const Cand* cands,
                                           const float* sw, int item, int total, const WarpCtx& w,
                                           int sub, ItemLoad<VT, NV>& ld) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr bool WIDE = (LANES == 32);
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
  const int W = p.level_w[l], H = p.level_h[l];
  const float ix = to_pixel(to_grid<MODE>(c.u), static_cast<float>(W));
  const float iy = to_pixel(to_grid<MODE>(c.v), static_cast<float>(H));
  const Footprint f = footprint(ix, iy, W, H);
  ld.w00 = (1.f - f.tx) * (1.f - f.ty);
  ld.w01 = f.tx * (1.f - f.ty);
  ld.w10 = (1.f - f.tx) * f.ty;
  ld.w11 = f.tx * f.ty;
  ld.wt = active ? wt : 0.f;
  ld.inb = (f.in00 ? ld.w00 : 0.f) + (f.in01 ? ld.w01 : 0.f) + (f.in10 ? ld.w10 : 0.f) +
           (f.in11 ? ld.w11 : 0.f);
  const VT* base = static_cast<const VT*>(p.value[l]);
  const size_t img = static_cast<size_t>(w.b) * p.N + n;
  const size_t row0 = (img * H + f.y0) * W;
  const size_t choff = (WIDE ? 0 : static_cast<size_t>(w.h) * kHeadDim) + sub * VEC;
  const VT* p00 = base + (row0 + f.x0) * p.C + choff;
  const VT* p10 = p00 + static_cast<size_t>(W) * p.C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int o = j * LANES * VEC;
    ld.r00[j] = ldg_nc_v4(p00 + o, active & f.in00);
    ld.r01[j] = ldg_nc_v4(p00 + p.C + o, active & f.in01);
    ld.r10[j] = ldg_nc_v4(p10 + o, active & f.in10);
    ld.r11[j] = ldg_nc_v4(p10 + p.C + o, active & f.in11);
  }
}

template <typename VT, int NV>
__device__ __forceinline__ void item_pin(ItemLoad<VT, NV>& ld) {
#pragma unroll
  for (int j = 0; j < NV; ++j) pin(ld.r00[j], ld.r01[j], ld.r10[j], ld.r11[j]);
}

template <typename VT, int NV>
__device__ __forceinline__ void item_consume(const ItemLoad<VT, NV>& ld,
                                             float (&acc)[ItemLoad<VT, NV>::PL], float& wsum) {
  constexpr int VEC = Slice<VT>::VEC;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float c00[VEC], c01[VEC], c10[VEC], c11[VEC];
    Slice<VT>::unpack(ld.r00[j], c00);
    Slice<VT>::unpack(ld.r01[j], c01);
    Slice<VT>::unpack(ld.r10[j], c10);
    Slice<VT>::unpack(ld.r11[j], c11);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float s = ld.w00 * c00[i] + ld.w01 * c01[i] + ld.w10 * c10[i] + ld.w11 * c11[i];
      acc[j * VEC + i] = fmaf(ld.wt, s, acc[j * VEC + i]);
    }
  }
  wsum = fmaf(ld.wt, ld.inb, wsum);
}

template <int MODE, typename VT, int LANES, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (LANES == 32) ? 2 : 4)
xview_fwd_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int PL = VEC * NV;
  constexpr int GROUPS = 32 / LANES;  // items per warp-wide load
  constexpr bool WIDE = (LANES == 32);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane_ = threadIdx.x & 31;
  const int grp = lane_ / LANES;
  const int sub = lane_ % LANES;
  const size_t warp_bytes = sizeof(float) * kMaxLP + sizeof(Cand) * cand_cap;
  float* sw = reinterpret_cast<float*>(smem_raw + warp * warp_bytes);
  Cand* cands = reinterpret_cast<Cand*>(sw + kMaxLP);

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    if (MODE == GD4D_MODE_C) head_softmax(p, w, sw);
    const int nvalid = build_candidates<MODE, Cand>(p, w, cands, p.mask != nullptr);

    float acc[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) acc[i] = 0.f;
    float wsum = 0.f;
    const int total = nvalid * p.L;
    for (int it0 = 0; it0 < total; it0 += 2 * GROUPS) {
      ItemLoad<VT, NV> la, lb;
      item_issue<MODE, VT, LANES, NV>(p, cands, sw, it0 + grp, total, w, sub, la);
      item_issue<MODE, VT, LANES, NV>(p, cands, sw, it0 + GROUPS + grp, total, w, sub, lb);
      item_pin<VT, NV>(la);
      item_pin<VT, NV>(lb);
      item_consume<VT, NV>(la, acc, wsum);
      item_consume<VT, NV>(lb, acc, wsum);
    }

    // ---- reduce across lane groups, store ----------------------------------------------
#pragma unroll
    for (int o = LANES; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < PL; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (grp == 0) {
      float* o = WIDE ? p.out + (static_cast<size_t>(w.bq) * p.Hh + w.h) * p.C
                      : p.out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
          *reinterpret_cast<float4*>(o + (j * LANES + sub) * VEC + i) = make_float4(
              acc[j * VEC + i], acc[j * VEC + i + 1], acc[j * VEC + i + 2], acc[j * VEC + i + 3]);
    }
    if (WIDE && p.wsum != nullptr && w.lane == 0) p.wsum[static_cast<size_t>(w.bq) * p.Hh + w.h] = wsum;
    __syncwarp();  // the per-warp shared-memory lists are reused by the next work item
  }
  work_end(p, wi);
}

template <int MODE, typename VT, int LANES, int NV>
static int launch_fwd(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_fwd_kernel<MODE, VT, LANES, NV>;
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem);
    if (e != cudaSuccess) return GD4D_ERR_CUDA;
  }
  int grid = g.grid;
  if (p.sched != nullptr) {  // persistent grid: one resident wave, warps claim work dynamically
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, g.block, g.smem) != cudaSuccess)
      return GD4D_ERR_CUDA;
    const long long resident = static_cast<long long>(sms) * (occ > 0 ? occ : 1);
    if (resident < grid) grid = static_cast<int>(resident);
  }
  kern<<<grid, g.block, g.smem, stream>>>(p, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int dispatch_forward(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  const bool bf16 = p.value_dtype == GD4D_BF16;
  if (p.mode == GD4D_MODE_A) {
    return bf16 ? launch_fwd<GD4D_MODE_A, __nv_bfloat16, 4, 1>(p, g, stream)
                : launch_fwd<GD4D_MODE_A, float, 8, 1>(p, g, stream);
  }
  if (p.mode == GD4D_MODE_C && !p.wide) {
    return bf16 ? launch_fwd<GD4D_MODE_C, __nv_bfloat16, 4, 1>(p, g, stream)
                : launch_fwd<GD4D_MODE_C, float, 8, 1>(p, g, stream);
  }
  if (p.mode == GD4D_MODE_C && p.wide) {
    if (bf16) return g.nv == 1 ? launch_fwd<GD4D_MODE_C, __nv_bfloat16, 32, 1>(p, g, stream)
                               : launch_fwd<GD4D_MODE_C, __nv_bfloat16, 32, 2>(p, g, stream);
    return g.nv == 1 ? launch_fwd<GD4D_MODE_C, float, 32, 1>(p, g, stream)
                     : launch_fwd<GD4D_MODE_C, float, 32, 2>(p, g, stream);
  }
  return GD4D_ERR_UNSUPPORTED;
}

}  // namespace gd4d
