// Wide-mode forward with a TMA staging path (Blackwell/Hopper bulk-copy engine).
//
// Same work decomposition and records as xview_fwd.cu, but the corner rows are not
// gathered into registers: for every batch of B items one elected lane issues
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` (SASS: UBLKCP)
// for the 4*B contiguous corner rows (C channels = 512 B or 1 KB each) into a per-warp
// shared-memory stage, and the warp waits on that stage's mbarrier before consuming it
// with conflict-free LDS.128.  S stages per warp keep up to S*B*4 rows in flight per warp
// without holding them in registers (the LDG kernel needs 64+ registers of landing space
// for 16 gathers), and the copy engine keeps fetching batch i+1.. while the warp does the
// FMAs of batch i.  Producer and consumer are the SAME warp, so a stage is free again
// after the __syncwarp() that follows its consumption; no empty-barrier is needed.
//
// r1 MEASUREMENT (B200, N=6 fp32, 1 KB rows): 67 us vs 48 us for the register-gather
// kernel, for every (B, S, warps/CTA) geometry tried (2x3x4, 4x2x4, 2x4x4, 2x2x8, 1x4x8,
// 2x3x8) -- the bulk-copy engine's per-operation cost dominates at 0.5-1 KB per copy.  The
// path is therefore OPT-IN (GD4D_FLAG_TMA_FORWARD); it stays built and parity-tested
// because wider rows (C >= 512 fp32) amortise the per-op cost.
#include "xview_common.cuh"
#include "xview_records.cuh"

namespace gd4d {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

template <typename VT, int NV, int B, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
xview_fwd_tma_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int PL = VEC * NV;
  constexpr int ROW = NV * 512;                    // bytes of one corner row (C channels)
  constexpr int STAGE = B * 4 * ROW;
  constexpr int MODE = GD4D_MODE_C;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t warp_bytes = static_cast<size_t>(S) * STAGE + sizeof(RecF) * 32 + sizeof(float) * kMaxLP +
                            16 * ((S + 1) / 2) + sizeof(Cand) * cand_cap;
  unsigned char* base = smem_raw + warp * ((warp_bytes + 127) / 128 * 128);
  unsigned char* stages = base;
  RecF* recs = reinterpret_cast<RecF*>(stages + S * STAGE);
  float* sw = reinterpret_cast<float*>(recs + 32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sw + kMaxLP);
  Cand* cands = reinterpret_cast<Cand*>(reinterpret_cast<unsigned char*>(bars) + 16 * ((S + 1) / 2));

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0;                              // bit s = parity to wait for on stage s
  const uint32_t stage0 = smem_u32(stages);

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    head_softmax(p, w, sw);
    const int nvalid = build_candidates<MODE, Cand>(p, w, cands, p.mask != nullptr);
    float acc[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) acc[i] = 0.f;
    float wsum_lane = 0.f;
    const int total = nvalid * p.L;
    for (int c0 = 0; c0 < total; c0 += 32) {
      float wt_item;
      recs[lane] = build_record<MODE, VT, true>(p, cands, sw, c0 + lane, total, w, wsum_lane, wt_item);
      __syncwarp();
      const int nchunk = min(32, total - c0);
      const int nb = (nchunk + B - 1) / B;
      auto issue = [&](int bi) {
        if (lane == 0) {
          const int s = bi % S;
          const uint32_t bar = smem_u32(bars + s);
          const uint32_t dst = stage0 + s * STAGE;
          mbar_arrive_expect_tx(bar, STAGE);
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const RecF r = recs[bi * B + u];       // slots past nchunk hold valid pointers, zero weights
            bulk_g2s(dst + (u * 4 + 0) * ROW, r.p00, ROW, bar);
            bulk_g2s(dst + (u * 4 + 1) * ROW, r.p01, ROW, bar);
            bulk_g2s(dst + (u * 4 + 2) * ROW, r.p10, ROW, bar);
            bulk_g2s(dst + (u * 4 + 3) * ROW, r.p11, ROW, bar);
          }
        }
      };
      // prologue: fill S-1 stages
#pragma unroll
      for (int s = 0; s < S - 1; ++s)
        if (s < nb) issue(s);
      for (int bi = 0; bi < nb; ++bi) {
        if (bi + S - 1 < nb) issue(bi + S - 1);    // that stage was consumed (and syncwarp'ed) at bi-1
        const int s = bi % S;
        mbar_wait(smem_u32(bars + s), (phase >> s) & 1u);
        phase ^= 1u << s;
        const unsigned char* st = stages + s * STAGE + lane * 16;
#pragma unroll
        for (int u = 0; u < B; ++u) {
          const RecF r = recs[bi * B + u];
          const float wgt[4] = {r.w00, r.w01, r.w10, r.w11};
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            float cc[4][VEC];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 t = *reinterpret_cast<const uint4*>(st + (u * 4 + c) * ROW + j * 512);
              Slice<VT>::unpack(t, cc[c]);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              float a = acc[j * VEC + i];
              a = fmaf(wgt[0], cc[0][i], a);
              a = fmaf(wgt[1], cc[1][i], a);
              a = fmaf(wgt[2], cc[2][i], a);
              a = fmaf(wgt[3], cc[3][i], a);
              acc[j * VEC + i] = a;
            }
          }
        }
        __syncwarp();                              // stage s (and, after the last batch, recs) reusable
      }
    }
    float* o = p.out + ((static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q) * p.C;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < VEC; i += 4)
        *reinterpret_cast<float4*>(o + (j * 32 + lane) * VEC + i) =
            make_float4(acc[j * VEC + i], acc[j * VEC + i + 1], acc[j * VEC + i + 2], acc[j * VEC + i + 3]);
    if (p.wsum != nullptr) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) wsum_lane += __shfl_xor_sync(0xffffffffu, wsum_lane, off);
      if (lane == 0) p.wsum[(static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q] = wsum_lane;
    }
    __syncwarp();
  }
  work_end(p, wi, WARPS);
}

template <typename VT, int NV, int B, int S, int WARPS>
static int launch_tma(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_fwd_tma_kernel<VT, NV, B, S, WARPS>;
  const size_t warp_bytes = static_cast<size_t>(S) * B * 4 * NV * 512 + sizeof(RecF) * 32 +
                            sizeof(float) * kMaxLP + 16 * ((S + 1) / 2) + sizeof(Cand) * g.cand_cap;
  const int smem = static_cast<int>(((warp_bytes + 127) / 128 * 128) * WARPS);
  if (smem > 227 * 1024) return GD4D_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return GD4D_ERR_CUDA;
  int dev = 0, sms = 0, occ = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem) != cudaSuccess)
    return GD4D_ERR_CUDA;
  const long long items = static_cast<long long>(p.B) * p.Q * p.Hh;
  long long grid = static_cast<long long>(sms) * (occ > 0 ? occ : 1);
  const long long need = (items + WARPS - 1) / WARPS;
  if (need < grid) grid = need;
  kern<<<static_cast<int>(grid), WARPS * 32, smem, stream>>>(p, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

// wide mode C, dynamic schedule only (the caller checks p.sched and the flag)
int dispatch_forward_tma(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  const bool bf16 = p.value_dtype == GD4D_BF16;
  if (bf16) return g.nv == 1 ? launch_tma<__nv_bfloat16, 1, 4, 3, 4>(p, g, stream)
                             : launch_tma<__nv_bfloat16, 2, 2, 3, 4>(p, g, stream);
  return g.nv == 1 ? launch_tma<float, 1, 4, 3, 4>(p, g, stream) : launch_tma<float, 2, 2, 3, 4>(p, g, stream);
}

}  // namespace gd4d
