// Probe of the tcgen05 no-swizzle operand layouts (dev tool): B's shared memory holds the ramp
// smem_float[i] = i, A is a K-major one-hot selector (A[m][k] = (m == k)), so after ONE
// tcgen05.mma (M=128, N=32, K=8, tf32) D[k][n] is the float index the tensor core fetched for B(k, n).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ int g_layout = 0;   // UMMA::LayoutType of the ramp operand (0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__global__ void probe(float* out, int b_mn, uint32_t lbo, uint32_t sbo, int a_mn, uint32_t albo, uint32_t asbo, int ramp_on_a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* sa = reinterpret_cast<float*>(smem);            // 16 KB
  float* sb = reinterpret_cast<float*>(smem + 16384);    // 16 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4096; i += 128) { sa[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  if (!ramp_on_a) {
    // A K-major canonical (LBO 128, SBO 256 for K = 8: 2 chunks): A[m][k] = (m == k), m < 8
    for (int i = tid; i < 8; i += 128) { const int m = i, k = i; sa[(m / 8) * 64 + (k / 4) * 32 + (m % 8) * 4 + (k % 4)] = 1.f; }
    for (int i = tid; i < 2048; i += 128) sb[i] = (float)i;
  } else {
    // B K-major canonical one-hot: B[n][k] = (n == k), n < 8;  A = ramp
    for (int i = tid; i < 8; i += 128) { const int n = i, k = i; sb[(n / 8) * 64 + (k / 4) * 32 + (n % 8) * 4 + (k % 4)] = 1.f; }
    for (int i = tid; i < 2048; i += 128) sa[i] = (float)i;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = ramp_on_a ? make_desc(smem_u32(sa), albo, asbo, g_layout) : make_desc(smem_u32(sa), 128, 256);
    const uint64_t db = ramp_on_a ? make_desc(smem_u32(sb), 128, 256) : make_desc(smem_u32(sb), lbo, sbo, g_layout);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (done) break;
    if (spin > (1u << 24)) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(r[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

static void run(const char* name, int b_mn, uint32_t lbo, uint32_t sbo, int a_mn, uint32_t albo, uint32_t asbo, int ramp_on_a, int layout = 0) {
  cudaMemcpyToSymbol(g_layout, &layout, sizeof(int));
  float* d; cudaMalloc(&d, 128 * 32 * 4); cudaMemset(d, 0, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(d, b_mn, lbo, sbo, a_mn, albo, asbo, ramp_on_a);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("== %s: %s\n", name, cudaGetErrorString(e));
  if (!ramp_on_a) {            // D[k][n] = index fetched for B(k, n): print k = 0..7, n = 0..31
    for (int k = 0; k < 8; ++k) { printf("k=%d:", k); for (int n = 0; n < 32; ++n) printf(" %4.0f", h[k * 32 + n]); printf("\n"); }
  } else {                     // D[m][k] = index fetched for A(m, k): print m = 0..15 and 64..67, k = 0..7
    for (int m = 0; m < 128; ++m) if (m < 10 || (m >= 30 && m < 36) || (m >= 64 && m < 67)) { printf("m=%d:", m); for (int k = 0; k < 8; ++k) printf(" %4.0f", h[m * 32 + k]); printf("\n"); }
  }
  cudaFree(d);
}

int main() {
  run("B K-major  lbo=128 sbo=256 (known good)", 0, 128, 256, 0, 0, 0, 0);
  run("B MN-major lbo=1024 sbo=128", 1, 1024, 128, 0, 0, 0, 0);
  run("B MN-major SW128 lbo=1024 sbo=1024", 1, 1024, 1024, 0, 0, 0, 0, 2);
  run("B MN-major SW128 lbo=2048 sbo=4096", 1, 2048, 4096, 0, 0, 0, 0, 2);
  run("B K-major  SW128 lbo=16 sbo=1024", 0, 16, 1024, 0, 0, 0, 0, 2);
  run("A MN-major SW128 lbo=1024 sbo=4096", 0, 0, 0, 1, 1024, 4096, 1, 2);
  run("A MN-major SW128 lbo=2048 sbo=1024", 0, 0, 0, 1, 2048, 1024, 1, 2);
  run("B MN-major SW32  lbo=1024 sbo=256", 1, 1024, 256, 0, 0, 0, 0, 6);
  return 0;
}
