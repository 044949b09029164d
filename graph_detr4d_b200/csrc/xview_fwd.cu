// Fused project -> mask -> bilinear-sample -> weight -> reduce forward kernel.
//
// One warp per (batch b, query q, head h) work item (claimed dynamically from a
// persistent grid when p.sched is set).  Three phases, all warp-synchronous:
//
//  1. candidates  the 32 lanes project the warp's N*P candidate points (camera n,
//                 graph-offset point p) in parallel, in the reference's exact op order
//                 (bit-exact mask), and ballot-compact the valid ones (~18 %) into a
//                 per-warp shared-memory list.
//  2. records     the (valid candidate, level) items are turned, 32 at a time and ONE
//                 LANE PER ITEM, into 48-byte records {4 corner pointers, 4 weights}.
//                 Out-of-map corners keep a valid (clamped) pointer and get weight 0,
//                 so the gather needs no predicates and no zero-fill.  (r1 finding: with
//                 every lane redoing the per-item scalar math the kernel issued ~220
//                 warp-instructions per item and was ISSUE-bound at 60 % issue-active;
//                 records cut that to ~60.)
//  3. gather      a lane GROUP owns an item and covers its channel run with 16-byte
//                 loads: narrow (mmcv layout, projected value, 32-ch head slice) 8 lanes
//                 fp32 / 4 lanes bf16; wide (gather-then-project, all C raw channels per
//                 head) 32 lanes x NV vectors.  16 independent 16-byte gathers per lane
//                 are issued back-to-back (volatile PTX + register pin + launch bounds)
//                 before the first FMA consumes one.  fp32 accumulate in registers, a
//                 shuffle reduce across groups, coalesced 16-byte stores.
//
// Replaces (reference, projects/mmdet3d_plugin/models/utils/):
//   mode A  detr3d_transformer.py:376-383 + feature_sampling :397-438
//   mode C  deform3d_cross_attn.py:227-258, 274, 281-284, 301-304, 320-324
#include "xview_common.cuh"
#include "xview_records.cuh"
#include "xview_sorted_ws.cuh"

namespace gd4d {

// EMIT (mode C wide only, GD4D_FLAG_FWD_EMIT): the lane that builds an item's gather record also writes the item's
// four corner contributions for the sorted backward (xview_bwd_sorted.cu) -- it holds exactly what that backward's
// emit kernel would recompute (corner rows, wt * bilinear weights) -- and takes their ranks in the row histogram.
template <int MODE, typename VT, int LANES, int NV, bool EMIT = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (LANES == 32) ? 2 : 3)
xview_fwd_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap, const __grid_constant__ SortedWs ews) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int PL = VEC * NV;
  constexpr int GROUPS = 32 / LANES;  // items per warp-wide load
  constexpr bool WIDE = (LANES == 32);
  constexpr int INF = 4 / NV;         // items in flight per group: 16 gathers per lane

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LANES;
  const int sub = lane % LANES;
  const size_t warp_bytes = sizeof(RecF) * 32 + sizeof(float) * kMaxLP + sizeof(Cand) * cand_cap;
  RecF* recs = reinterpret_cast<RecF*>(smem_raw + warp * warp_bytes);
  float* sw = reinterpret_cast<float*>(recs + 32);
  Cand* cands = reinterpret_cast<Cand*>(sw + kMaxLP);
  const int lane_off = sub * 16;      // this lane's 16 bytes inside a (LANES*16)-byte run
  const bool l2_prefetch = (p.flags & GD4D_FLAG_L2_PREFETCH) != 0;

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    if (MODE == GD4D_MODE_C) head_softmax(p, w, sw);
    const int nvalid = build_candidates<MODE, Cand>(p, w, cands, p.mask != nullptr);

    float acc[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) acc[i] = 0.f;
    float wsum_lane = 0.f;
    const int total = nvalid * p.L;
    int emit_base = 0;
    if (EMIT) {
      unsigned v = 0;
      if (lane == 0 && total > 0) v = atomicAdd(ews.counters, static_cast<unsigned>(4 * total));
      emit_base = static_cast<int>(__shfl_sync(0xffffffffu, v, 0));
      if (lane == 0) ews.base[w.bq * p.Hh + w.h] = emit_base;
    }
    for (int c0 = 0; c0 < total; c0 += 32) {
      float wt_item;
      const RecF rec = build_record<MODE, VT, WIDE>(p, cands, sw, c0 + lane, total, w, wsum_lane, wt_item);
      recs[lane] = rec;
      if (EMIT && c0 + lane < total) {
        const int l = (c0 + lane) % p.L;
        const char* vbase = static_cast<const char*>(p.value[l]);
        const char* zrow = reinterpret_cast<const char*>(g_zero_row);
        const long long rowb = static_cast<long long>(p.C) * sizeof(VT);
        const int go_row = (w.b * p.Hh + w.h) * p.Q + w.q;
        const char* ptr[4] = {rec.p00, rec.p01, rec.p10, rec.p11};
        const float cw[4] = {rec.w00, rec.w01, rec.w10, rec.w11};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool in_map = ptr[j] != zrow;
          const long long rl = in_map ? (ptr[j] - vbase) / rowb : 0;       // pixel row inside the level
          emit_contribution(ews, emit_base + (c0 + lane) * 4 + j, in_map, ptr[j], rl * p.C, l,
                            static_cast<int>(ews.level_row0[l] + rl), go_row, cw[j]);
        }
      }
      if (MODE != GD4D_MODE_C) sw[lane] = wt_item;   // mode A: applied AFTER nan_to_num(sample); sw is free (no softmax)
      __syncwarp();
      const int nchunk = min(32, total - c0);
      for (int j0 = 0; j0 < nchunk; j0 += INF * GROUPS) {
        uint4 raw[INF][4][NV];
        float wgt[INF][4];
        float wtA[INF];
#pragma unroll
        for (int u = 0; u < INF; ++u) {
          const RecF r = recs[j0 + u * GROUPS + grp];
          wgt[u][0] = r.w00; wgt[u][1] = r.w01; wgt[u][2] = r.w10; wgt[u][3] = r.w11;
          wtA[u] = (MODE != GD4D_MODE_C) ? sw[j0 + u * GROUPS + grp] : 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const int o = lane_off + j * LANES * 16;
            raw[u][0][j] = ldg_nc_v4_all(r.p00 + o);
            raw[u][1][j] = ldg_nc_v4_all(r.p01 + o);
            raw[u][2][j] = ldg_nc_v4_all(r.p10 + o);
            raw[u][3][j] = ldg_nc_v4_all(r.p11 + o);
          }
        }
        if (l2_prefetch) {  // next batch's corner rows -> L2 while this batch's gathers are in flight
#pragma unroll
          for (int u = 0; u < INF; ++u) {
            const int it = j0 + (INF + u) * GROUPS + grp;
            if (it < nchunk) {
              const RecF* r = recs + it;
#pragma unroll
              for (int j = 0; j < NV; ++j) {
                const int o = lane_off + j * LANES * 16;
                prefetch_l2(r->p00 + o); prefetch_l2(r->p01 + o);
                prefetch_l2(r->p10 + o); prefetch_l2(r->p11 + o);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < INF; ++u)
#pragma unroll
          for (int j = 0; j < NV; ++j) pin(raw[u][0][j], raw[u][1][j], raw[u][2][j], raw[u][3][j]);
#pragma unroll
        for (int u = 0; u < INF; ++u)
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            float c00[VEC], c01[VEC], c10[VEC], c11[VEC];
            Slice<VT>::unpack(raw[u][0][j], c00);
            Slice<VT>::unpack(raw[u][1][j], c01);
            Slice<VT>::unpack(raw[u][2][j], c10);
            Slice<VT>::unpack(raw[u][3][j], c11);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              if (MODE == GD4D_MODE_C) {
                float a = acc[j * VEC + i];
                a = fmaf(wgt[u][0], c00[i], a);
                a = fmaf(wgt[u][1], c01[i], a);
                a = fmaf(wgt[u][2], c10[i], a);
                a = fmaf(wgt[u][3], c11[i], a);
                acc[j * VEC + i] = a;
              } else {  // detr3d_transformer.py:378: nan_to_num on the bilinear sample, then the weight
                float sm = wgt[u][0] * c00[i];
                sm = fmaf(wgt[u][1], c01[i], sm);
                sm = fmaf(wgt[u][2], c10[i], sm);
                sm = fmaf(wgt[u][3], c11[i], sm);
                acc[j * VEC + i] = fmaf(wtA[u], nan_to_num_(sm), acc[j * VEC + i]);
              }
            }
          }
      }
      __syncwarp();  // records are rebuilt by the next chunk
    }

    // ---- reduce across lane groups, store ----------------------------------------------
#pragma unroll
    for (int o = LANES; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < PL; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (grp == 0) {
      float* o = WIDE ? p.out + ((static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q) * p.C
                      : p.out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
          *reinterpret_cast<float4*>(o + (j * LANES + sub) * VEC + i) = make_float4(
              acc[j * VEC + i], acc[j * VEC + i + 1], acc[j * VEC + i + 2], acc[j * VEC + i + 3]);
    }
    if (WIDE && p.wsum != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wsum_lane += __shfl_xor_sync(0xffffffffu, wsum_lane, o);
      if (lane == 0) p.wsum[(static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q] = wsum_lane;
    }
    __syncwarp();  // the per-warp shared-memory lists are reused by the next work item
  }
  work_end(p, wi);
}

template <int MODE, typename VT, int LANES, int NV, bool EMIT = false>
static int launch_fwd(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_fwd_kernel<MODE, VT, LANES, NV, EMIT>;
  SortedWs ews{};
  if (EMIT) {
    if ((reinterpret_cast<uintptr_t>(p.bwd_ws) & 255u) != 0) return GD4D_ERR_ALIGN;
    const long long need = sorted_ws_layout(p, &ews, static_cast<char*>(p.bwd_ws));
    if (need < 0 || p.bwd_ws_bytes < need) return GD4D_ERR_DIMS;
  }
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
  kern<<<grid, g.block, g.smem, stream>>>(p, g.cand_cap, ews);
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
    if ((p.flags & GD4D_FLAG_FWD_EMIT) && p.bwd_ws != nullptr) {
      if (bf16) return g.nv == 1 ? launch_fwd<GD4D_MODE_C, __nv_bfloat16, 32, 1, true>(p, g, stream)
                                 : launch_fwd<GD4D_MODE_C, __nv_bfloat16, 32, 2, true>(p, g, stream);
      return g.nv == 1 ? launch_fwd<GD4D_MODE_C, float, 32, 1, true>(p, g, stream)
                       : launch_fwd<GD4D_MODE_C, float, 32, 2, true>(p, g, stream);
    }
    if (bf16) return g.nv == 1 ? launch_fwd<GD4D_MODE_C, __nv_bfloat16, 32, 1>(p, g, stream)
                               : launch_fwd<GD4D_MODE_C, __nv_bfloat16, 32, 2>(p, g, stream);
    return g.nv == 1 ? launch_fwd<GD4D_MODE_C, float, 32, 1>(p, g, stream)
                     : launch_fwd<GD4D_MODE_C, float, 32, 2>(p, g, stream);
  }
  return GD4D_ERR_UNSUPPORTED;
}

}  // namespace gd4d
