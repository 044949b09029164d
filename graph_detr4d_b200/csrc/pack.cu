// NCHW -> channel-last packing of the FPN feature maps, with optional bf16 cast.
//
// The reference re-does this layout change in EVERY decoder layer
// (deform3d_cross_attn.py:264-269: view/flatten/transpose + cat, then value_proj
// materialises it again); here it runs once per forward and the six layers share
// the packed maps.  Pure HBM-bound transpose: 32x32 tiles through padded shared
// memory, coalesced 128-byte reads along W*H and coalesced writes along C.
#include "xview_common.cuh"

namespace gd4d {

template <typename TI, typename TO>
__device__ __forceinline__ TO cvt(TI v);
template <> __device__ __forceinline__ float cvt<float, float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<float, __nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16, __nv_bfloat16>(__nv_bfloat16 v) {
  return v;
}

// grid: (ceil(HW/32), ceil(C/32), images), block (32, 8)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) pack_nchw_kernel(const TI* __restrict__ src, TO* __restrict__ dst,
                                                        int C, int HW) {
  __shared__ TI tile[32][33];
  const size_t img = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TI* s = src + img * static_cast<size_t>(C) * HW;
  TO* d = dst + img * static_cast<size_t>(C) * HW;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) tile[j][threadIdx.x] = s[static_cast<size_t>(c) * HW + hw];
  }
  __syncthreads();
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int hw = hw0 + j, c = c0 + threadIdx.x;
    if (c < C && hw < HW) d[static_cast<size_t>(hw) * C + c] = cvt<TI, TO>(tile[threadIdx.x][j]);
  }
}

// Vectorised variant (fp32 source, HW % 4 == 0, C % 4 == 0): 32-channel x 128-pixel tiles.
// A thread loads four float4 (4 consecutive pixels of 4 consecutive channels: each warp-wide
// load is 512 contiguous bytes of one channel row), transposes the 4x4 block in registers and
// stores four float4 (4 channels of one pixel) into a [128][32]-float tile whose 16-byte
// column index is XOR-swizzled with the pixel quad, so both the STS.128 and the LDS.128 of the
// write-out phase are bank-conflict free.  Write-out: 8 lanes cover the 128 bytes (32 channels)
// of one pixel, a warp 4 pixels per instruction.
template <typename TO>
__device__ __forceinline__ void store4(TO* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<unsigned*>(&a);
  u.y = *reinterpret_cast<unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// four consecutive source elements as float4 (fp32: one 16-byte load, bf16: one 8-byte load)
template <typename TI>
__device__ __forceinline__ float4 load4(const TI* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) pack_nchw_v4_kernel(const TI* __restrict__ src, TO* __restrict__ dst,
                                                           int C, int HW) {
  __shared__ __align__(16) float tile[128 * 32];
  const size_t img = blockIdx.z;
  const int hw0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
  const TI* s = src + img * static_cast<size_t>(C) * HW;
  TO* d = dst + img * static_cast<size_t>(C) * HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const int c = c0 + warp * 4, hw = hw0 + lane * 4;
    float4 r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + k < C && hw < HW) r[k] = load4<TI>(s + static_cast<size_t>(c + k) * HW + hw);
    }
    const int q = (warp ^ (lane & 7)) * 4;         // swizzled 16-byte column of channels c..c+3
    float* t = tile + (lane * 4) * 32 + q;
    *reinterpret_cast<float4*>(t) = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
    *reinterpret_cast<float4*>(t + 32) = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
    *reinterpret_cast<float4*>(t + 64) = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
    *reinterpret_cast<float4*>(t + 96) = make_float4(r[0].w, r[1].w, r[2].w, r[3].w);
  }
  __syncthreads();
  const int cq = lane & 7;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int hwl = it * 32 + warp * 4 + (lane >> 3);
    const int hw = hw0 + hwl, c = c0 + cq * 4;
    if (hw < HW && c < C) {
      const float4 v = *reinterpret_cast<const float4*>(tile + hwl * 32 + ((cq ^ ((hwl >> 2) & 7)) * 4));
      store4<TO>(d + static_cast<size_t>(hw) * C + c, v);
    }
  }
}

int dispatch_pack(const void* src, void* dst, int src_dtype, int dst_dtype, int64_t images, int C,
                  int H, int W, cudaStream_t stream) {
  const int HW = H * W;
  const bool v4_types = (src_dtype == GD4D_F32 && (dst_dtype == GD4D_F32 || dst_dtype == GD4D_BF16)) ||
                        (src_dtype == GD4D_BF16 && dst_dtype == GD4D_BF16);
  if (v4_types && HW % 4 == 0 && C % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    for (int64_t i0 = 0; i0 < images; i0 += 65535) {
      const int64_t cnt = images - i0 < 65535 ? images - i0 : 65535;
      dim3 grid((HW + 127) / 128, (C + 31) / 32, static_cast<unsigned>(cnt));
      const size_t off = static_cast<size_t>(i0) * C * HW;
      if (src_dtype == GD4D_BF16)
        pack_nchw_v4_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(
            static_cast<const __nv_bfloat16*>(src) + off, static_cast<__nv_bfloat16*>(dst) + off, C, HW);
      else if (dst_dtype == GD4D_F32)
        pack_nchw_v4_kernel<float, float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src) + off,
                                                                    static_cast<float*>(dst) + off, C, HW);
      else
        pack_nchw_v4_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(
            static_cast<const float*>(src) + off, static_cast<__nv_bfloat16*>(dst) + off, C, HW);
      if (cudaGetLastError() != cudaSuccess) return GD4D_ERR_CUDA;
    }
    return GD4D_OK;
  }
  dim3 block(32, 8);
  // gridDim.z is limited to 65535 images per launch
  for (int64_t i0 = 0; i0 < images; i0 += 65535) {
    const int64_t cnt = images - i0 < 65535 ? images - i0 : 65535;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, static_cast<unsigned>(cnt));
    const size_t off = static_cast<size_t>(i0) * C * HW;
    if (src_dtype == GD4D_F32 && dst_dtype == GD4D_F32)
      pack_nchw_kernel<float, float><<<grid, block, 0, stream>>>(
          static_cast<const float*>(src) + off, static_cast<float*>(dst) + off, C, HW);
    else if (src_dtype == GD4D_F32 && dst_dtype == GD4D_BF16)
      pack_nchw_kernel<float, __nv_bfloat16><<<grid, block, 0, stream>>>(
          static_cast<const float*>(src) + off, static_cast<__nv_bfloat16*>(dst) + off, C, HW);
    else if (src_dtype == GD4D_BF16 && dst_dtype == GD4D_BF16)
      pack_nchw_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(src) + off, static_cast<__nv_bfloat16*>(dst) + off, C, HW);
    else
      return GD4D_ERR_UNSUPPORTED;
    if (cudaGetLastError() != cudaSuccess) return GD4D_ERR_CUDA;
  }
  return GD4D_OK;
}

}  // namespace gd4d
