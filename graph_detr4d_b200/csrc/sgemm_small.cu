// Exact-fp32 GEMM for the decoder's small-M shapes (include/gd4d_glue.h).
//
//   C[M,N] = A . B^T (+ bias) (relu),  M = B*Q = 900 rows, K and N in {256, 512}: 118 MFLOP per call, ~190 calls
//   per training step.  cuBLAS serves them with cutlass3x_sm100_simt_sgemm 32x32x16 at 7-12 us per call
//   (~16 TFLOP/s, 53 % of the r1 step).  The tensor-core alternative (csrc/gemm_tf32x3.cu, error-compensated
//   3xTF32) is NOT faster here: tools/umma_latency.cu measures 153 cycles per tcgen05.mma.kind::tf32
//   (M = 128, K = 8) for every N from 32 to 256, so a 128-row tile needs K/8 * 3 * 153 cycles = 7.5 us at
//   K = 256 however narrow it is, and 900 rows are only 8 such tiles.  This kernel keeps the FFMA pipe busy:
//
//   * CTA tile 32 x 64, 256 threads = two k-halves of 128 threads (intra-CTA split-K: each half owns 16 of the
//     32 k of every chunk; the halves are summed through shared memory in the epilogue), 4 x 4 outputs per thread
//   * 4-stage cp.async ring of operand tiles (16-byte copies, zero-filled outside the matrix)
//   * both operand majors without a transpose pass.  K-major source (row = m|n, contiguous = k): tile kept
//     [row][k] with a 144-byte row pitch, a thread owns rows t, t+T, t+2T, t+3T and reads float4 along k
//     (conflict-free: 9 x 16 B pitch).  MN-major source (row = k, contiguous = m|n): tile kept [k][row], a thread
//     owns 4 consecutive rows and reads one float4 per k.  Either way 8 LDS.128 feed 64 FFMA.
//   * epilogue: + bias, relu, coalesced stores
// Same arithmetic class as the library call it replaces (fp32 FMA chains; only the summation order differs).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gd4d_glue.h"

namespace gd4d {
namespace sg {

constexpr int kBM = 32, kBN = 64, kBK = 32, kThreads = 256, kStages = 4;

struct Args {
  const float* A; const float* B; float* C; const float* bias;
  long long lda, ldb, ldc, sA, sB, sC;
  int M, N, K, relu;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// One operand tile: R rows (m | n) x 32 k
template <int R, bool MN>
struct Tile {
  static constexpr int kPitch = MN ? R : 36;                    // floats per shared-memory row
  static constexpr int kFloats = MN ? kBK * kPitch : R * kPitch;
  static constexpr int kChunks = R * 8;                          // 16-byte chunks

  __device__ __forceinline__ static void copy(float* dst, const float* src, long long ld, int r0, int k0, int r_valid,
                                              int k_valid, int tid) {
    const uint32_t base = smem_u32(dst);
    for (int c = tid; c < kChunks; c += kThreads) {
      int row, col;
      uint32_t off;
      bool ok;
      if (!MN) {
        const int r = c >> 3, kc = c & 7;
        row = r0 + r; col = k0 + kc * 4;
        ok = row < r_valid && col + 3 < k_valid;
        off = (r * kPitch + kc * 4) * 4;
      } else {
        const int k = c / (R / 4), rc = c % (R / 4);
        row = k0 + k; col = r0 + rc * 4;
        ok = row < k_valid && col + 3 < r_valid;
        off = (k * kPitch + rc * 4) * 4;
      }
      const float* g = ok ? src + static_cast<long long>(row) * ld + col : src;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + off), "l"(g), "r"(ok ? 16u : 0u) : "memory");
    }
  }

  // fragment of one 4-k step: f[i][kk] = X(row_i, k4*4 + kk), rows t + T*i (K-major) or 4t + i (MN-major)
  template <int T>
  __device__ __forceinline__ static void frag(const float* tile, int t, int k4, float (&f)[4][4]) {
    if (!MN) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(tile + (t + T * i) * kPitch + k4 * 4);
        f[i][0] = v.x; f[i][1] = v.y; f[i][2] = v.z; f[i][3] = v.w;
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 v = *reinterpret_cast<const float4*>(tile + (k4 * 4 + kk) * kPitch + 4 * t);
        f[0][kk] = v.x; f[1][kk] = v.y; f[2][kk] = v.z; f[3][kk] = v.w;
      }
    }
  }
  template <int T>
  __device__ __forceinline__ static int row_of(int t, int i) { return MN ? 4 * t + i : t + T * i; }
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads) sgemm_small_kernel(const Args g) {
  using TA = Tile<kBM, A_MN>;
  using TB = Tile<kBN, B_MN>;
  constexpr int kStageFloats = TA::kFloats + TB::kFloats;
  extern __shared__ __align__(16) float smem[];

  const int tid = threadIdx.x;
  const int half = tid >> 7;                     // which 16 of the chunk's 32 k this thread accumulates
  const int t = tid & 127, tx = t & 15, ty = t >> 4;   // 16 thread columns (n) x 8 thread rows (m)
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const float* A = g.A + blockIdx.z * g.sA;
  const float* B = g.B + blockIdx.z * g.sB;
  float* C = g.C + blockIdx.z * g.sC;
  const int nchunks = (g.K + kBK - 1) / kBK;

  auto issue = [&](int kb) {
    if (kb < nchunks) {
      float* st = smem + (kb % kStages) * kStageFloats;
      TA::copy(st, A, g.lda, m0, kb * kBK, g.M, g.K, tid);
      TB::copy(st + TA::kFloats, B, g.ldb, n0, kb * kBK, g.N, g.K, tid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kb = 0; kb < nchunks; ++kb) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
    __syncthreads();                              // chunk kb is visible; everyone is done with chunk kb-1's stage
    issue(kb + kStages - 1);                      // refill the stage chunk kb-1 used
    const float* st = smem + (kb % kStages) * kStageFloats;
#pragma unroll
    for (int q = 0; q < kBK / 8; ++q) {           // this half's 4 of the chunk's 8 four-k steps
      const int k4 = half * (kBK / 8) + q;
      float a[4][4], b[4][4];
      TA::template frag<8>(st, ty, k4, a);
      TB::template frag<16>(st + TA::kFloats, tx, k4, b);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], b[j][kk], acc[i][j]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();                                // the ring is free: reuse it for the k-half reduction
  float* red = smem;                              // [128 threads][16] floats, thread-major (conflict-free float4)
  if (half == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(red + (i * 128 + t) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  if (half == 1) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 o = *reinterpret_cast<const float4*>(red + (i * 128 + t) * 4);
    acc[i][0] += o.x; acc[i][1] += o.y; acc[i][2] += o.z; acc[i][3] += o.w;
  }
  const bool vec_ok = B_MN && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + TA::template row_of<8>(ty, i);
    if (m >= g.M) continue;
    float* crow = C + static_cast<long long>(m) * g.ldc;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + TB::template row_of<16>(tx, j);
      float x = acc[i][j];
      if (g.bias != nullptr && n < g.N) x += __ldg(g.bias + n);
      if (g.relu) x = fmaxf(x, 0.f);
      o[j] = x;
    }
    const int nfirst = n0 + TB::template row_of<16>(tx, 0);
    if (vec_ok && nfirst + 3 < g.N) {             // MN-major B: the thread's 4 columns are consecutive
      *reinterpret_cast<float4*>(crow + nfirst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + TB::template row_of<16>(tx, j);
        if (n < g.N) crow[n] = o[j];              // K-major B: lanes tx = 0..15 write 16 consecutive floats per j
      }
    }
  }
}

template <bool A_MN, bool B_MN>
static int launch(const Args& g, int batch, cudaStream_t stream) {
  using TA = Tile<kBM, A_MN>;
  using TB = Tile<kBN, B_MN>;
  constexpr int smem = kStages * (TA::kFloats + TB::kFloats) * 4;
  static_assert(smem >= 128 * 16 * 4, "the ring must hold the k-half reduction buffer");
  auto kern = sgemm_small_kernel<A_MN, B_MN>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return GD4D_ERR_CUDA;
  }
  dim3 grid((g.N + kBN - 1) / kBN, (g.M + kBM - 1) / kBM, batch);
  kern<<<grid, kThreads, smem, stream>>>(g);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

}  // namespace sg
}  // namespace gd4d

extern "C" int gd4d_sgemm_small(const float* A, int64_t lda, int32_t a_mn_major, const float* B, int64_t ldb,
                                int32_t b_mn_major, float* C, int64_t ldc, const float* bias, int32_t relu, int32_t M,
                                int32_t N, int32_t K, int32_t batch, int64_t stride_a, int64_t stride_b,
                                int64_t stride_c, void* cuda_stream) {
  if (A == nullptr || B == nullptr || C == nullptr) return GD4D_ERR_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || batch > 65535) return GD4D_ERR_DIMS;
  if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u) != 0) return GD4D_ERR_ALIGN;
  if (lda % 4 != 0 || ldb % 4 != 0 || stride_a % 4 != 0 || stride_b % 4 != 0) return GD4D_ERR_ALIGN;
  if ((a_mn_major ? M : K) % 4 != 0 || (b_mn_major ? N : K) % 4 != 0) return GD4D_ERR_UNSUPPORTED;
  if ((M + gd4d::sg::kBM - 1) / gd4d::sg::kBM > 65535) return GD4D_ERR_DIMS;
  gd4d::sg::Args g{A, B, C, bias, lda, ldb, ldc, stride_a, stride_b, stride_c, M, N, K, relu ? 1 : 0};
  auto st = static_cast<cudaStream_t>(cuda_stream);
  switch ((a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0)) {
    case 0: return gd4d::sg::launch<false, false>(g, batch, st);
    case 1: return gd4d::sg::launch<false, true>(g, batch, st);
    case 2: return gd4d::sg::launch<true, false>(g, batch, st);
    default: return gd4d::sg::launch<true, true>(g, batch, st);
  }
}
