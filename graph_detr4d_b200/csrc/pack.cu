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

int dispatch_pack(const void* src, void* dst, int src_dtype, int dst_dtype, int64_t images, int C,
                  int H, int W, cudaStream_t stream) {
  const int HW = H * W;
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
