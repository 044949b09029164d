// l2_roofs.cu -- measured roofs of the access patterns the cross-view sampling kernels use
// (dev tool, standalone binary; prints one JSON object per line).
//
// The fused sampling kernels gather / scatter 0.5-1 KB channel-last pixel rows at data-dependent
// positions; ~60 % of those requests hit the 126 MB L2, so HBM bandwidth is not the ceiling that
// binds them.  This tool measures what is:
//   ldg_rows   random 1 KB row gathers with ld.global.nc.v4 (the kernels' LDG path)
//   red_rows   random 1 KB row reductions with red.global.add.v4.f32 (the backward's scatter)
//   tma_load   random 2x2-pixel x 256-channel boxes through a 4-D tensor map
//              (cp.async.bulk.tensor.4d -> UTMALDG), per-warp mbarrier ring
//   tma_red    the same boxes reduced into global memory by the TMA
//              (cp.reduce.async.bulk.tensor.4d .add -> UTMAREDG)
// over footprints that fit L2 (1 image of a 116x200x256 fp32 level = 23.75 MB), straddle it
// (6 images = the bench's level 0) and exceed it (24 images).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/l2_roofs tools/l2_roofs.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const char* p) {
  uint4 t;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
  return t;
}
__device__ __forceinline__ uint4 ldg_nc_na_v4(const char* p) {
  uint4 t;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
  return t;
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int ROW = 1024;  // bytes of one fp32 256-channel pixel row

// ---- 1. LDG row gathers: RIF rows (2 x 16 B per lane per row) in flight per warp -----------------
template <int RIF, bool NOALLOC>
__global__ void __launch_bounds__(256) ldg_rows(const char* __restrict__ buf, uint32_t nrows, int iters,
                                                 uint32_t* sink) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint4 v[RIF][2];
#pragma unroll
    for (int r = 0; r < RIF; ++r) {
      const uint32_t row = hash32((gw * iters + it) * RIF + r) % nrows;
      const char* p = buf + static_cast<size_t>(row) * ROW + lane * 16;
      v[r][0] = NOALLOC ? ldg_nc_na_v4(p) : ldg_nc_v4(p);
      v[r][1] = NOALLOC ? ldg_nc_na_v4(p + 512) : ldg_nc_v4(p + 512);
    }
#pragma unroll
    for (int r = 0; r < RIF; ++r) acc ^= v[r][0].x ^ v[r][0].w ^ v[r][1].y ^ v[r][1].z;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- 2. RED row scatters -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) red_rows(float* __restrict__ buf, uint32_t nrows, int iters) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    const uint32_t row = hash32(gw * iters + it) % nrows;
    float* p = buf + static_cast<size_t>(row) * (ROW / 4) + lane * 4;
    red_add_v4(p, 1.f, 1.f, 1.f, 1.f);
    red_add_v4(p + 128, 1.f, 1.f, 1.f, 1.f);
  }
}

// ---- 3. gather + scatter of the same rows (the backward's mix) ---------------------------------------
template <int RIF>
__global__ void __launch_bounds__(256) ldg_red_rows(const char* __restrict__ src, float* __restrict__ dst,
                                                     uint32_t nrows, int iters, uint32_t* sink) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint4 v[RIF][2];
    uint32_t rows[RIF];
#pragma unroll
    for (int r = 0; r < RIF; ++r) {
      rows[r] = hash32((gw * iters + it) * RIF + r) % nrows;
      const char* p = src + static_cast<size_t>(rows[r]) * ROW + lane * 16;
      v[r][0] = ldg_nc_v4(p);
      v[r][1] = ldg_nc_v4(p + 512);
    }
#pragma unroll
    for (int r = 0; r < RIF; ++r) {
      float* p = dst + static_cast<size_t>(rows[r]) * (ROW / 4) + lane * 4;
      red_add_v4(p, 1.f, 1.f, 1.f, 1.f);
      red_add_v4(p + 128, 1.f, 1.f, 1.f, 1.f);
      acc ^= v[r][0].x ^ v[r][0].w ^ v[r][1].y ^ v[r][1].z;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- TMA helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_red_add_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}

constexpr int BOX = 4 * ROW;  // 2x2 pixels x 256 fp32 channels

// ---- 4. TMA 2x2 box loads: per-warp ring of D stages ----------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) tma_load(const __grid_constant__ CUtensorMap tm, int W, int H, int IMG,
                                                 int iters, uint32_t* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * nw + warp;
  unsigned char* stages = smem + static_cast<size_t>(warp) * D * BOX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(nw) * D * BOX) + warp * D;
  if (lane == 0) {
    for (int s = 0; s < D; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int it) {
    const int s = it % D;
    const uint32_t h = hash32(gw * iters + it);
    const int x = static_cast<int>(h % (W + 1)) - 1;             // -1 .. W-1 (border boxes are zero-filled)
    const int y = static_cast<int>((h >> 10) % (H + 1)) - 1;
    const int img = static_cast<int>((h >> 20) % IMG);
    const uint32_t bar = smem_u32(bars + s);
    mbar_arrive_expect_tx(bar, BOX);
    tma_load_4d(smem_u32(stages + s * BOX), &tm, 0, x, y, img, bar);
  };
  if (lane == 0)
    for (int s = 0; s < D && s < iters; ++s) issue(s);
  uint32_t acc = 0, phase = 0;
  for (int it = 0; it < iters; ++it) {
    const int s = it % D;
    mbar_wait(smem_u32(bars + s), (phase >> s) & 1u);
    phase ^= 1u << s;
    const uint4* st = reinterpret_cast<const uint4*>(stages + s * BOX) + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint4 v = st[k * 32];
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    __syncwarp();
    if (lane == 0 && it + D < iters) issue(it + D);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- 5. TMA 2x2 box reductions (add.f32) ---------------------------------------------------------------
template <int D, bool POS>
__global__ void __launch_bounds__(256) tma_red(const __grid_constant__ CUtensorMap tm, int W, int H, int IMG,
                                                int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * nw + warp;
  unsigned char* stages = smem + static_cast<size_t>(warp) * D * BOX;
  for (int it = 0; it < iters; ++it) {
    const int s = it % D;
    if (it >= D) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(D - 1) : "memory");
      __syncwarp();
    }
    float4* st = reinterpret_cast<float4*>(stages + s * BOX) + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) st[k * 32] = make_float4(1.f, 1.f, 1.f, 1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const uint32_t h = hash32(gw * iters + it);
      const int x = POS ? static_cast<int>(h % (W - 1)) : static_cast<int>(h % (W + 1)) - 1;
      const int y = POS ? static_cast<int>((h >> 10) % (H - 1)) : static_cast<int>((h >> 10) % (H + 1)) - 1;
      const int img = static_cast<int>((h >> 20) % IMG);
      tma_red_add_4d(&tm, 0, x, y, img, smem_u32(stages + s * BOX));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}


// ---- 5b. non-tensor bulk reductions: cp.reduce.async.bulk .add.f32 of CONTIGUOUS runs (1 or 2 KB) --------------
__device__ __forceinline__ void bulk_red_add_f32(float* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
template <int D, int RUN>   // RUN bytes per op; a 4 KB stage = 4096/RUN ops at random rows
__global__ void __launch_bounds__(256) bulk_red(float* __restrict__ buf, uint32_t nrows, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * nw + warp;
  unsigned char* stages = smem + static_cast<size_t>(warp) * D * BOX;
  for (int it = 0; it < iters; ++it) {
    const int s = it % D;
    if (it >= D) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(D - 1) : "memory");
      __syncwarp();
    }
    float4* st = reinterpret_cast<float4*>(stages + s * BOX) + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) st[k * 32] = make_float4(1.f, 1.f, 1.f, 1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < BOX / RUN; ++r) {
        const uint32_t row = hash32((gw * iters + it) * 4 + r) % (nrows - 1);
        bulk_red_add_f32(buf + static_cast<size_t>(row) * (ROW / 4), smem_u32(stages + s * BOX + r * RUN), RUN);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}

// ---- 6. TMA box load + TMA box reduce of the same box (the backward's mix through the TMA) -----------------
template <int D>
__global__ void __launch_bounds__(256) tma_load_red(const __grid_constant__ CUtensorMap tsrc,
                                                     const __grid_constant__ CUtensorMap tdst, int W, int H, int IMG,
                                                     int iters, uint32_t* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t gw = blockIdx.x * nw + warp;
  unsigned char* stages = smem + static_cast<size_t>(warp) * 2 * D * BOX;       // D load + D reduce stages
  unsigned char* rstages = stages + D * BOX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(nw) * 2 * D * BOX) + warp * D;
  if (lane == 0) {
    for (int s = 0; s < D; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto coords = [&](int it, int& x, int& y, int& img) {
    const uint32_t h = hash32(gw * iters + it);
    x = static_cast<int>(h % (W + 1)) - 1;
    y = static_cast<int>((h >> 10) % (H + 1)) - 1;
    img = static_cast<int>((h >> 20) % IMG);
  };
  auto issue = [&](int it) {
    const int s = it % D;
    int x, y, img;
    coords(it, x, y, img);
    const uint32_t bar = smem_u32(bars + s);
    mbar_arrive_expect_tx(bar, BOX);
    tma_load_4d(smem_u32(stages + s * BOX), &tsrc, 0, x, y, img, bar);
  };
  if (lane == 0)
    for (int s = 0; s < D && s < iters; ++s) issue(s);
  uint32_t acc = 0, phase = 0;
  for (int it = 0; it < iters; ++it) {
    const int s = it % D;
    mbar_wait(smem_u32(bars + s), (phase >> s) & 1u);
    phase ^= 1u << s;
    if (it >= D) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(D - 1) : "memory");
      __syncwarp();
    }
    const uint4* st = reinterpret_cast<const uint4*>(stages + s * BOX) + lane;
    float4* rt = reinterpret_cast<float4*>(rstages + s * BOX) + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint4 v = st[k * 32];
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
      rt[k * 32] = make_float4(1.f, 1.f, 1.f, 1.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      int x, y, img;
      coords(it, x, y, img);
      tma_red_add_4d(&tdst, 0, x, y, img, smem_u32(rstages + s * BOX));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (it + D < iters) issue(it + D);
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- host ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (fn == nullptr || q != cudaDriverEntryPointSuccess) {
    fprintf(stderr, "cuTensorMapEncodeTiled unavailable\n");
    exit(1);
  }
  return reinterpret_cast<EncodeFn>(fn);
}

static CUtensorMap make_map(EncodeFn enc, void* base, int W, int H, int IMG, CUtensorMapL2promotion promo) {
  CUtensorMap m;
  const cuuint64_t dims[4] = {256, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)IMG};
  const cuuint64_t strides[3] = {(cuuint64_t)ROW, (cuuint64_t)W * ROW, (cuuint64_t)H * W * ROW};
  const cuuint32_t box[4] = {256, 2, 2, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(1);
  }
  return m;
}

template <typename F>
static float time_best(F launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

static void report(const char* test, const char* variant, int imgs, int ctas_per_sm, int warps, double bytes, float ms) {
  printf("{\"test\": \"%s\", \"variant\": \"%s\", \"footprint_mb\": %.1f, \"ctas_per_sm\": %d, \"warps_per_cta\": %d, "
         "\"bytes\": %.0f, \"ms\": %.4f, \"GBps\": %.1f}\n",
         test, variant, imgs * 116.0 * 200.0 * ROW / 1e6, ctas_per_sm, warps, bytes, ms, bytes / ms / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  // usage: l2_roofs <group> [imgs]   groups: ldg red mix tma_load tma_red tma_red_pos bulk_red1k bulk_red2k tma_load_red
  // one group per process: a faulting variant (sticky CUDA error) must not take the others down
  const char* group = argc > 1 ? argv[1] : "ldg";
  const int only_imgs = argc > 2 ? atoi(argv[2]) : 0;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int W = 200, H = 116, MAXIMG = 24;
  const size_t maxbytes = static_cast<size_t>(MAXIMG) * W * H * ROW;
  char* src;
  float* dst;
  uint32_t* sink;
  CK(cudaMalloc(&src, maxbytes));
  CK(cudaMalloc(&dst, maxbytes));
  CK(cudaMalloc(&sink, 16));
  CK(cudaMemset(src, 1, maxbytes));
  CK(cudaMemset(dst, 0, maxbytes));
  EncodeFn enc = get_encode();
  const double target_bytes = 4e9;   // per measurement
  auto is = [&](const char* g) { return strcmp(group, g) == 0; };

  for (int imgs : {1, 6, 24}) {
    if (only_imgs && imgs != only_imgs) continue;
    const uint32_t nrows = static_cast<uint32_t>(imgs) * W * H;
    for (int cps : {2, 4, 8}) {
      const int grid = sms * cps, warps = grid * 8;
      if (is("ldg")) {
        {
          const int iters = static_cast<int>(target_bytes / (double(warps) * 8 * ROW)) + 1;
          const double bytes = double(warps) * iters * 8 * ROW;
          report("ldg_rows", "rif8", imgs, cps, 8, bytes, time_best([&] { ldg_rows<8, false><<<grid, 256>>>(src, nrows, iters, sink); }));
          report("ldg_rows", "rif8_noalloc", imgs, cps, 8, bytes, time_best([&] { ldg_rows<8, true><<<grid, 256>>>(src, nrows, iters, sink); }));
        }
        {
          const int iters = static_cast<int>(target_bytes / (double(warps) * 4 * ROW)) + 1;
          const double bytes = double(warps) * iters * 4 * ROW;
          report("ldg_rows", "rif4", imgs, cps, 8, bytes, time_best([&] { ldg_rows<4, false><<<grid, 256>>>(src, nrows, iters, sink); }));
        }
      }
      if (is("red")) {
        const int iters = static_cast<int>(target_bytes / 2 / (double(warps) * ROW)) + 1;
        const double bytes = double(warps) * iters * ROW;
        report("red_rows", "v4", imgs, cps, 8, bytes, time_best([&] { red_rows<<<grid, 256>>>(dst, nrows, iters); }));
      }
      if (is("mix")) {
        const int iters = static_cast<int>(target_bytes / 2 / (double(warps) * 4 * ROW)) + 1;
        const double bytes = double(warps) * iters * 4 * ROW;   // bytes gathered (= bytes reduced)
        report("ldg_red_rows", "rif4", imgs, cps, 8, bytes, time_best([&] { ldg_red_rows<4><<<grid, 256>>>(src, dst, nrows, iters, sink); }));
      }
    }
    // TMA: 1 CTA per SM, 8 warps x D stages x 4 KB
    for (CUtensorMapL2promotion promo : {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B}) {
      const char* pn = promo == CU_TENSOR_MAP_L2_PROMOTION_NONE ? "" : "_promo256";
      CUtensorMap ms = make_map(enc, src, W, H, imgs, promo);
      CUtensorMap md = make_map(enc, dst, W, H, imgs, promo);
      const int grid = sms, warps = grid * 8;
      const int iters = static_cast<int>(target_bytes / (double(warps) * BOX)) + 1;
      const double bytes = double(warps) * iters * BOX;
      const int it2 = iters / 2 + 1;
      char name[64];
      if (is("tma_load")) {
        {
          constexpr int D = 6;
          const int smem = 8 * D * BOX + 8 * D * 8;
          CK(cudaFuncSetAttribute(tma_load<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          snprintf(name, sizeof name, "d6%s", pn);
          report("tma_load", name, imgs, 1, 8, bytes, time_best([&] { tma_load<D><<<grid, 256, smem>>>(ms, W, H, imgs, iters, sink); }));
        }
        {
          constexpr int D = 3;
          const int smem = 8 * D * BOX + 8 * D * 8;
          CK(cudaFuncSetAttribute(tma_load<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          snprintf(name, sizeof name, "d3%s", pn);
          report("tma_load", name, imgs, 1, 8, bytes, time_best([&] { tma_load<D><<<grid, 256, smem>>>(ms, W, H, imgs, iters, sink); }));
          snprintf(name, sizeof name, "d3x2cta%s", pn);   // 2 CTAs per SM
          report("tma_load", name, imgs, 2, 8, bytes * 2, time_best([&] { tma_load<D><<<grid * 2, 256, smem>>>(ms, W, H, imgs, iters, sink); }));
        }
      }
      if (is("tma_red") || is("tma_red_pos")) {
        constexpr int D = 4;
        const int smem = 8 * D * BOX;
        snprintf(name, sizeof name, "d4%s", pn);
        if (is("tma_red")) {
          CK(cudaFuncSetAttribute(tma_red<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          report("tma_red", name, imgs, 1, 8, double(warps) * it2 * BOX, time_best([&] { tma_red<D, false><<<grid, 256, smem>>>(md, W, H, imgs, it2); }));
        } else {
          CK(cudaFuncSetAttribute(tma_red<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          report("tma_red_pos", name, imgs, 1, 8, double(warps) * it2 * BOX, time_best([&] { tma_red<D, true><<<grid, 256, smem>>>(md, W, H, imgs, it2); }));
        }
      }
      if (is("tma_load_red")) {
        constexpr int D = 3;
        const int smem = 8 * 2 * D * BOX + 8 * D * 8;
        CK(cudaFuncSetAttribute(tma_load_red<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        snprintf(name, sizeof name, "d3%s", pn);
        report("tma_load_red", name, imgs, 1, 8, double(warps) * it2 * BOX,
               time_best([&] { tma_load_red<D><<<grid, 256, smem>>>(ms, md, W, H, imgs, it2, sink); }));
      }
      if (promo != CU_TENSOR_MAP_L2_PROMOTION_NONE) continue;
      if (is("bulk_red1k") || is("bulk_red2k")) {
        constexpr int D = 4;
        const int smem = 8 * D * BOX;
        if (is("bulk_red1k")) {
          CK(cudaFuncSetAttribute(bulk_red<D, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          report("bulk_red", "1k_d4", imgs, 1, 8, double(warps) * it2 * BOX, time_best([&] { bulk_red<D, 1024><<<grid, 256, smem>>>(dst, nrows, it2); }));
        } else {
          CK(cudaFuncSetAttribute(bulk_red<D, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          report("bulk_red", "2k_d4", imgs, 1, 8, double(warps) * it2 * BOX, time_best([&] { bulk_red<D, 2048><<<grid, 256, smem>>>(dst, nrows, it2); }));
        }
      }
    }
  }
  return 0;
}
