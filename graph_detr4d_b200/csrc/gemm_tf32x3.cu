// fp32-accurate small-M GEMM on the 5th-generation tensor cores (include/gd4d_glue.h).
//
// The decoder around the sampling kernels is ~190 GEMMs per step with M = B*Q = 900 rows and K, N in
// {256, 512}: 118 MFLOP each.  cuBLAS runs them as SIMT fp32 (cutlass3x_sm100_simt_sgemm 32x32x16, 10-26 us
// each, 53 % of the r1 step) because TF32 would break the 1e-5 parity the fp32 reference demands.  Here:
//
//   C[M,N] = A . B^T (+ bias) (relu)         fp32 in, fp32 out, error-compensated 3xTF32:
//       x = hi + lo,  hi = x & 0xffffe000 (exact tf32),  lo = x - hi (exact, 13 significant bits)
//       A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo   (the dropped lo.lo term is 2^-22 relative)
//   three tcgen05.mma.kind::tf32 per 8-wide k step, fp32 accumulators in tensor memory.
//
//   * CTA tile 128 x BN (BN = 32 | 64) x 32, 128 threads, one elected thread issues the MMAs
//   * global -> shared: cp.async 16-byte copies into a 4-deep ring of RAW tiles (the whole first 128 k of both
//     operands is in flight before the first MMA), zero-filled outside the matrix
//   * split pass (all threads): raw tile -> (hi, lo) tiles in the canonical NO-SWIZZLE K-major core-matrix
//     layout (8 rows x 16 B core matrices of 128 B; the 8 k-chunks of a row group contiguous), two stages.
//     MN-major sources (row = k, contiguous = m|n: W in dY.W, dY and X in dY^T.X) are TRANSPOSED here --
//     measured on this part (tools/umma_probe.cu): tcgen05.mma.kind::tf32 returns zeros for MN-major operand
//     descriptors in every layout, while K-major works with and without swizzle.
//   * the tensor core's fp32 accumulate truncates (measured 3.4e-6 relative at K = 256 with one accumulator,
//     growing with K): the hi.hi products rotate over three accumulators and the small terms get a fourth,
//     summed in round-to-nearest fp32 in the epilogue
//   * tcgen05.commit -> mbarrier tells the split pass a stage is free again
//   * epilogue: tcgen05.ld 32x32b (warp w owns accumulator lanes 32w..32w+31 = rows), + bias, relu, 16-byte stores
// Every mbarrier wait is a BOUNDED spin that traps: a malformed descriptor faults instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gd4d_glue.h"

namespace gd4d {

constexpr int kBM = 128, kBK = 32, kThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();   // never completes: fail loudly instead of hanging the device
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout NONE [61,64)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

struct GemmArgs {
  const float* A; const float* B; float* C; const float* bias;
  long long lda, ldb, ldc, sA, sB, sC;
  int M, N, K, relu;
};

constexpr int kRawStages = 4;

// One operand tile of R rows (m | n) x 32 k.
//   K-major source  (MN = false): src[r][k], k contiguous.  Raw copy: [R][32] floats, row pitch 144 B (the 16 B pad
//                   keeps the split pass's quarter-warp loads -- 8 rows, same k chunk -- on distinct banks).
//   MN-major source (MN = true) : src[k][r], r contiguous.  Raw copy: [32][R] floats, row pitch 4R B.
// Canonical (hi / lo) tile: offset(r, k) = (r / 8) * 1024 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4.
template <int R, bool MN>
struct Operand {
  static constexpr int kPitch = MN ? R * 4 : 144;
  static constexpr int kRawBytes = MN ? 32 * kPitch : R * kPitch;
  static constexpr int kTileBytes = R * 32 * 4;
  static constexpr int kChunks = R * 8;                      // 16-byte chunks per tile
  static_assert(kChunks % kThreads == 0, "tile does not split over the threads");

  // cp.async of one raw tile; (r0, k0) = tile origin; rows >= r_valid or k >= k_valid are zero-filled
  __device__ __forceinline__ static void copy(unsigned char* raw, const float* src, long long ld, int r0, int k0,
                                              int r_valid, int k_valid, int tid) {
    const uint32_t base = smem_u32(raw);
#pragma unroll
    for (int it = 0; it < kChunks / kThreads; ++it) {
      const int c = it * kThreads + tid;
      int row, col;                                           // source row / first source column of this chunk
      uint32_t off;
      bool ok;
      if (!MN) {
        const int r = c >> 3, kc = c & 7;                     // 8 consecutive threads: one row's 128 B
        row = r0 + r; col = k0 + kc * 4;
        ok = row < r_valid && col + 3 < k_valid;
        off = r * kPitch + kc * 16;
      } else {
        const int k = c / (R / 4), rc = c % (R / 4);          // R/4 consecutive threads: one k row's 4R B
        row = k0 + k; col = r0 + rc * 4;
        ok = row < k_valid && col + 3 < r_valid;
        off = k * kPitch + rc * 16;
      }
      const float* g = ok ? src + static_cast<long long>(row) * ld + col : src;
      const uint32_t nbytes = ok ? 16u : 0u;                  // src-size 0: the 16 bytes are zero-filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + off), "l"(g), "r"(nbytes) : "memory");
    }
  }

  // raw tile -> canonical hi / lo tiles
  __device__ __forceinline__ static void split(const unsigned char* raw, unsigned char* hi, unsigned char* lo, int tid) {
#pragma unroll
    for (int it = 0; it < kChunks / kThreads; ++it) {
      const int c = it * kThreads + tid;
      float4 v;
      int r, kc;
      if (!MN) {
        r = (c & 7) + ((c >> 6) << 3);                        // quarter-warp: 8 rows of one k chunk
        kc = (c >> 3) & 7;
        v = *reinterpret_cast<const float4*>(raw + r * kPitch + kc * 16);
      } else {
        r = c % R;                                            // a warp: 32 consecutive m|n of 4 consecutive k
        kc = c / R;
        const float* p = reinterpret_cast<const float*>(raw + (kc * 4) * kPitch) + r;
        v.x = p[0]; v.y = p[kPitch / 4]; v.z = p[2 * (kPitch / 4)]; v.w = p[3 * (kPitch / 4)];
      }
      const int off = (r >> 3) * 1024 + kc * 128 + (r & 7) * 16;
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  }
};

template <int BN, bool A_MN, bool B_MN>
struct GemmSmem {
  using OpA = Operand<kBM, A_MN>;
  using OpB = Operand<BN, B_MN>;
  static constexpr int kRaw = OpA::kRawBytes + OpB::kRawBytes;            // one raw stage
  static constexpr int kSplit = 2 * OpA::kTileBytes + 2 * OpB::kTileBytes;  // one split stage: Ahi Alo Bhi Blo
  static constexpr int kTotal = kRawStages * kRaw + 2 * kSplit;
};

// A_MN / B_MN: the SOURCE operand is MN-major (rows = k, contiguous = m | n); the tensor core always sees K-major
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads) gemm_tf32x3_kernel(const GemmArgs g) {
  using S = GemmSmem<BN, A_MN, B_MN>;
  using OpA = typename S::OpA;
  using OpB = typename S::OpB;
  constexpr int kAccs = 4, kCols = kAccs * BN;                            // tensor-memory columns (power of 2 >= 32)
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_free[2];
  __shared__ uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;
  unsigned char* raw0 = smem;
  unsigned char* split0 = smem + kRawStages * S::kRaw;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const float* A = g.A + blockIdx.z * g.sA;
  const float* B = g.B + blockIdx.z * g.sB;
  float* C = g.C + blockIdx.z * g.sC;
  const int nchunks = (g.K + kBK - 1) / kBK;

  auto issue = [&](int kb) {                                  // one cp.async group per chunk (empty past the end)
    if (kb < nchunks) {
      unsigned char* raw = raw0 + (kb % kRawStages) * S::kRaw;
      OpA::copy(raw, A, g.lda, m0, kb * kBK, g.M, g.K, tid);
      OpB::copy(raw + OpA::kRawBytes, B, g.ldb, n0, kb * kBK, g.N, g.K, tid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int s = 0; s < kRawStages; ++s) issue(s);              // the memory system is busy before anything else

  if (tid == 0) {
    mbar_init(&bar_free[0], 1); mbar_init(&bar_free[1], 1); mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();                                  // .sync.aligned below needs the warp converged after the tid == 0 branch
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32 K-major, N >> 3, M >> 4
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                             (static_cast<uint32_t>(kBM >> 4) << 24);
  // K-major no-swizzle descriptors: LBO (between the two 16-byte k chunks of one MMA) = 128,
  // SBO (between 8-row groups) = 1024; one k step (8 floats) = 2 chunks = 256 B further
  const int ksteps = nchunks * (kBK / 8);
  const int rot = ksteps >= 3 ? 3 : 1;                        // accumulators the hi.hi products rotate over

  for (int kb = 0; kb < nchunks; ++kb) {
    const int s = kb & 1;
    unsigned char* raw = raw0 + (kb % kRawStages) * S::kRaw;
    unsigned char* st = split0 + s * S::kSplit;
    asm volatile("cp.async.wait_group %0;" ::"n"(kRawStages - 1) : "memory");   // this thread's copies of chunk kb landed
    __syncthreads();                                                              // ... and everyone else's
    if (kb >= 2) mbar_wait(&bar_free[s], ((kb >> 1) - 1) & 1);   // the MMAs that read this split stage have retired
    OpA::split(raw, st, st + OpA::kTileBytes, tid);
    OpB::split(raw + OpA::kRawBytes, st + 2 * OpA::kTileBytes, st + 2 * OpA::kTileBytes + OpB::kTileBytes, tid);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();                                               // also: the raw stage is free again
    issue(kb + kRawStages);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + OpA::kTileBytes, b_hi = a_hi + 2 * OpA::kTileBytes,
                     b_lo = b_hi + OpB::kTileBytes;
#pragma unroll
      for (int ks = 0; ks < kBK / 8; ++ks) {
        const int kk = kb * (kBK / 8) + ks;
        const uint64_t dah = make_desc(a_hi + ks * 256, 128, 1024), dal = make_desc(a_lo + ks * 256, 128, 1024);
        const uint64_t dbh = make_desc(b_hi + ks * 256, 128, 1024), dbl = make_desc(b_lo + ks * 256, 128, 1024);
        mma_tf32(tmem + 3 * BN, dal, dbh, idesc, kk != 0);         // small terms: their own accumulator
        mma_tf32(tmem + 3 * BN, dah, dbl, idesc, true);
        mma_tf32(tmem + (kk % rot) * BN, dah, dbh, idesc, kk >= rot);
      }
      mma_commit(&bar_free[s]);
      if (kb + 1 == nchunks) mma_commit(&bar_acc);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  mbar_wait(&bar_acc, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w owns accumulator lanes (= tile rows) 32w .. 32w+31; 16 columns of the 4 accumulators at a time
  const int row = m0 + warp * 32 + lane;
  const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int a = kAccs - 1; a >= 0; --a) {                          // small terms first
      if (a < 3 && a >= rot) continue;                              // never written for very short K
      uint32_t r[16];
      const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16) + a * BN + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = (a == kAccs - 1) ? __uint_as_float(r[e]) : acc[e] + __uint_as_float(r[e]);
    }
    if (row < g.M) {
      float* crow = C + static_cast<long long>(row) * g.ldc;
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const int n = n0 + c0 + c;
        if (n >= g.N) break;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = acc[c + e];
          if (g.bias != nullptr && n + e < g.N) x += __ldg(g.bias + n + e);
          if (g.relu) x = fmaxf(x, 0.f);
          o[e] = x;
        }
        if (vec_ok && n + 3 < g.N) {
          *reinterpret_cast<float4*>(crow + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < g.N) crow[n + e] = o[e];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols) : "memory");
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const GemmArgs& g, int batch, cudaStream_t stream) {
  constexpr int smem = GemmSmem<BN, A_MN, B_MN>::kTotal;
  auto kern = gemm_tf32x3_kernel<BN, A_MN, B_MN>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return GD4D_ERR_CUDA;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + kBM - 1) / kBM, batch);
  kern<<<grid, kThreads, smem, stream>>>(g);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

}  // namespace gd4d

extern "C" int gd4d_gemm_tf32x3(const float* A, int64_t lda, int32_t a_mn_major, const float* B, int64_t ldb,
                                int32_t b_mn_major, float* C, int64_t ldc, const float* bias, int32_t relu, int32_t M,
                                int32_t N, int32_t K, int32_t batch, int64_t stride_a, int64_t stride_b,
                                int64_t stride_c, void* cuda_stream) {
  if (A == nullptr || B == nullptr || C == nullptr) return GD4D_ERR_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || batch > 65535) return GD4D_ERR_DIMS;
  if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u) != 0) return GD4D_ERR_ALIGN;
  // 16-byte operand chunks: the contiguous extent and every row start must be multiples of 4 floats
  if (lda % 4 != 0 || ldb % 4 != 0 || stride_a % 4 != 0 || stride_b % 4 != 0) return GD4D_ERR_ALIGN;
  if ((a_mn_major ? M : K) % 4 != 0 || (b_mn_major ? N : K) % 4 != 0) return GD4D_ERR_UNSUPPORTED;
  if ((M + gd4d::kBM - 1) / gd4d::kBM > 65535) return GD4D_ERR_DIMS;
  gd4d::GemmArgs g{A, B, C, bias, lda, ldb, ldc, stride_a, stride_b, stride_c, M, N, K, relu ? 1 : 0};
  auto st = static_cast<cudaStream_t>(cuda_stream);
  // enough CTAs to cover the machine: 128 x 32 tiles unless 128 x 64 tiles already give >= 96 of them
  const long long ctas64 = static_cast<long long>((N + 63) / 64) * ((M + gd4d::kBM - 1) / gd4d::kBM) * batch;
  const bool bn64 = N > 32 && ctas64 >= 96;
  const int key = (bn64 ? 4 : 0) | (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
  switch (key) {
    case 0: return gd4d::launch_gemm<32, false, false>(g, batch, st);
    case 1: return gd4d::launch_gemm<32, false, true>(g, batch, st);
    case 2: return gd4d::launch_gemm<32, true, false>(g, batch, st);
    case 3: return gd4d::launch_gemm<32, true, true>(g, batch, st);
    case 4: return gd4d::launch_gemm<64, false, false>(g, batch, st);
    case 5: return gd4d::launch_gemm<64, false, true>(g, batch, st);
    case 6: return gd4d::launch_gemm<64, true, false>(g, batch, st);
    default: return gd4d::launch_gemm<64, true, true>(g, batch, st);
  }
}
