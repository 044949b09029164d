// tcgen05.mma.kind::tf32 issue/latency microbenchmark (dev tool): one CTA, one thread issues `count` MMAs
// (M = 128, N = n, K = 8) on zeroed K-major operands, rotating over `accs` accumulators, then commits and
// waits.  Reports cycles per MMA: same-accumulator chains vs independent accumulators, small vs large N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_latency tools/umma_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void bench(long long* out, int n, int accs, int count, int distinct_ops) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384;
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t ks = distinct_ops ? (i & 3) * 256 : 0;
      const uint64_t da = make_desc(a0 + ks, 128, 1024), db = make_desc(b0 + ks, 128, 1024);
      const uint32_t d = tmem + (i % accs) * n;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(i >= accs ? 1u : 0u) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    for (uint32_t spin = 0;; ++spin) {
      uint32_t done;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      if (done) break;
      if (spin > (1u << 26)) __trap();
    }
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int count = 512;
  const int cfgs[][3] = {{32, 1, 0}, {32, 2, 0}, {32, 4, 0}, {32, 8, 0}, {32, 16, 0}, {64, 1, 0}, {64, 4, 0}, {64, 8, 0},
                         {128, 1, 0}, {128, 4, 0}, {256, 1, 0}, {256, 2, 0}, {32, 1, 1}, {32, 8, 1}, {256, 1, 1}};
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<1, 128, 48 * 1024>>>(d, c[0], c[1], count, c[2]);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (rep == 1) printf("N=%3d accs=%2d distinct_operands=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", c[0], c[1], c[2],
                           (double)h[0] / count, (double)h[1] / count, cudaGetErrorString(e));
    }
  }
  return 0;
}
