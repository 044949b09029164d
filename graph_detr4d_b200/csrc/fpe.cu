// Feature position-embedding block of the PE head (include/gd4d_fpe.h):
// level padding masks, the 3-D sine embedding computed straight from the image sizes, and the
// SE-gate / add combine with its backward.  HBM-write-bound elementwise kernels; consecutive threads
// own consecutive pixels of one channel plane so every warp store is one contiguous 128-byte line.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gd4d_fpe.h"

namespace gd4d {

// F.interpolate(mode='nearest') source index (ATen nearest_neighbor_compute_source_index with
// scale = float(in) / out): min(int(floorf(dst * scale)), in - 1)
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  return min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(dst), scale))), in_size - 1);
}

__global__ void __launch_bounds__(256) level_mask_kernel(const int32_t* __restrict__ img_hw,
                                                          uint8_t* __restrict__ mask, int H, int W, int pad_h,
                                                          int pad_w) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int bn = blockIdx.y;
  const int y = pix / W, x = pix - y * W;
  const int ih = __ldg(img_hw + 2 * bn), iw = __ldg(img_hw + 2 * bn + 1);
  const bool valid = (nearest_src(y, pad_h, H) < ih) & (nearest_src(x, pad_w, W) < iw);   // :529-531, :535-536
  mask[static_cast<size_t>(bn) * H * W + pix] = valid ? 0 : 1;
}

// rows / columns of the level that map inside the camera's image: the source index is monotone in the
// destination index, so the valid set is a prefix and not_mask is a rectangle [0,hv) x [0,wv)
__device__ __forceinline__ int valid_prefix(int size, int pad, int img) {
  int lo = 0, hi = size;                                   // first dst with src(dst) >= img
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nearest_src(mid, pad, size) < img) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kMaxCams = 64;

__global__ void __launch_bounds__(128) sine_pe3d_kernel(const int32_t* __restrict__ img_hw,
                                                         const float* __restrict__ dim_t, float* __restrict__ out,
                                                         int N, int H, int W, int pad_h, int pad_w, int F,
                                                         int normalize, float scale, float eps, float offset) {
  __shared__ int s_hv[kMaxCams], s_wv[kMaxCams];
  extern __shared__ float s_dim[];                          // F floats
  const int b = blockIdx.z, n = blockIdx.y;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    s_hv[i] = valid_prefix(H, pad_h, __ldg(img_hw + 2 * (b * N + i)));
    s_wv[i] = valid_prefix(W, pad_w, __ldg(img_hw + 2 * (b * N + i) + 1));
  }
  for (int i = threadIdx.x; i < F; i += blockDim.x) s_dim[i] = __ldg(dim_t + i);
  __syncthreads();
  const int HW = H * W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const int y = pix / W, x = pix - y * W;
  const int hv = s_hv[n], wv = s_wv[n];
  // cumsums of not_mask (positional_encoding.py:71-74), closed form
  int cn = 0, cn_last = 0;
  for (int i = 0; i < N; ++i) {
    const int v = (y < s_hv[i]) & (x < s_wv[i]);
    cn_last += v;
    cn += (i <= n) ? v : 0;
  }
  float en = static_cast<float>(cn);
  float ey = (x < wv) ? static_cast<float>(min(y + 1, hv)) : 0.f;
  float ex = (y < hv) ? static_cast<float>(min(x + 1, wv)) : 0.f;
  if (normalize) {                                          // :75-81   (e + offset) / (e_last + eps) * scale
    const float ln = static_cast<float>(cn_last);
    const float ly = (x < wv) ? static_cast<float>(hv) : 0.f;
    const float lx = (y < hv) ? static_cast<float>(wv) : 0.f;
    en = __fmul_rn(__fdiv_rn(__fadd_rn(en, offset), __fadd_rn(ln, eps)), scale);
    ey = __fmul_rn(__fdiv_rn(__fadd_rn(ey, offset), __fadd_rn(ly, eps)), scale);
    ex = __fmul_rn(__fdiv_rn(__fadd_rn(ex, offset), __fadd_rn(lx, eps)), scale);
  }
  float* o = out + (static_cast<size_t>(b * N + n) * 3 * F) * HW + pix;
  const float e3[3] = {en, ey, ex};                         // cat((pos_n, pos_y, pos_x)) :99
  // stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).view(B,N,H,W,-1) on the 5-D p stacks BEFORE the
  // feature axis (:90-98): channels [0, F/2) are sin(p[2j]), channels [F/2, F) are cos(p[2j+1]) -- two
  // halves, not the interleaving the 4-D mmdet original produces.
  const int Fh = F >> 1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float* ok = o + static_cast<size_t>(k) * F * HW;
    for (int j = 0; j < Fh; ++j) {
      const float d0 = s_dim[2 * j], d1 = s_dim[2 * j + 1];
      const float a0 = __fdiv_rn(e3[k], d0);                // :85-87
      float sn, cs;
      if (d0 == d1) {
        sincosf(a0, &sn, &cs);
      } else {
        sn = sinf(a0);
        cs = cosf(__fdiv_rn(e3[k], d1));
      }
      ok[static_cast<size_t>(j) * HW] = sn;
      ok[static_cast<size_t>(Fh + j) * HW] = cs;
    }
  }
}

__device__ __forceinline__ float sigmoid_(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

__global__ void __launch_bounds__(256) fpe_combine_fwd_kernel(const float4* __restrict__ feat,
                                                               const float4* __restrict__ pe,
                                                               const float4* __restrict__ gate,
                                                               const float4* __restrict__ sine,
                                                               float4* __restrict__ out, int64_t n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 f = feat[i], p = pe[i], g = gate[i], s = sine[i];
  float4 o;   // feat + ((pe * sigmoid(gate)) + sine)      :243, :552-553
  o.x = __fadd_rn(f.x, __fadd_rn(__fmul_rn(p.x, sigmoid_(g.x)), s.x));
  o.y = __fadd_rn(f.y, __fadd_rn(__fmul_rn(p.y, sigmoid_(g.y)), s.y));
  o.z = __fadd_rn(f.z, __fadd_rn(__fmul_rn(p.z, sigmoid_(g.z)), s.z));
  o.w = __fadd_rn(f.w, __fadd_rn(__fmul_rn(p.w, sigmoid_(g.w)), s.w));
  out[i] = o;
}

__global__ void __launch_bounds__(256) fpe_combine_fwd_tail(const float* feat, const float* pe, const float* gate,
                                                             const float* sine, float* out, int64_t start,
                                                             int64_t n) {
  const int64_t i = start + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(feat[i], __fadd_rn(__fmul_rn(pe[i], sigmoid_(gate[i])), sine[i]));
}

__global__ void __launch_bounds__(256) fpe_combine_bwd_kernel(const float* __restrict__ go,
                                                               const float* __restrict__ pe,
                                                               const float* __restrict__ gate,
                                                               float* __restrict__ gpe, float* __restrict__ ggate,
                                                               int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = go[i];
  const float s = sigmoid_(gate[i]);
  if (gpe != nullptr) gpe[i] = g * s;
  if (ggate != nullptr) ggate[i] = (g * pe[i]) * (s * (1.f - s));
}

}  // namespace gd4d

extern "C" {

int gd4d_level_mask(const int32_t* img_hw, uint8_t* mask, int32_t BN, int32_t H, int32_t W, int32_t pad_h,
                    int32_t pad_w, void* cuda_stream) {
  if (img_hw == nullptr || mask == nullptr) return GD4D_ERR_NULL;
  if (BN <= 0 || BN > 65535 || H <= 0 || W <= 0 || pad_h <= 0 || pad_w <= 0) return GD4D_ERR_DIMS;
  dim3 grid((H * W + 255) / 256, BN);
  gd4d::level_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(img_hw, mask, H, W, pad_h, pad_w);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int gd4d_sine_pe3d(const int32_t* img_hw, const float* dim_t, float* out, int32_t B, int32_t N, int32_t H,
                   int32_t W, int32_t pad_h, int32_t pad_w, int32_t F, int32_t normalize, float scale, float eps,
                   float offset, void* cuda_stream) {
  if (img_hw == nullptr || dim_t == nullptr || out == nullptr) return GD4D_ERR_NULL;
  if (B <= 0 || B > 65535 || N <= 0 || N > gd4d::kMaxCams || H <= 0 || W <= 0 || pad_h <= 0 || pad_w <= 0 ||
      F <= 0 || F > 4096 || (F & 1))          // odd num_feats cannot be stacked in the reference either
    return GD4D_ERR_DIMS;
  dim3 grid((H * W + 127) / 128, N, B);
  gd4d::sine_pe3d_kernel<<<grid, 128, F * sizeof(float), static_cast<cudaStream_t>(cuda_stream)>>>(
      img_hw, dim_t, out, N, H, W, pad_h, pad_w, F, normalize, scale, eps, offset);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int gd4d_fpe_combine_fwd(const float* feat, const float* pe, const float* gate, const float* sine, float* out,
                         int64_t n, void* cuda_stream) {
  if (feat == nullptr || pe == nullptr || gate == nullptr || sine == nullptr || out == nullptr) return GD4D_ERR_NULL;
  if (n <= 0) return GD4D_ERR_DIMS;
  auto st = static_cast<cudaStream_t>(cuda_stream);
  const bool al = ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(pe) |
                    reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(sine) |
                    reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  const int64_t n4 = al ? n / 4 : 0;
  if (n4 > 0)
    gd4d::fpe_combine_fwd_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(feat), reinterpret_cast<const float4*>(pe),
        reinterpret_cast<const float4*>(gate), reinterpret_cast<const float4*>(sine),
        reinterpret_cast<float4*>(out), n4);
  const int64_t rest = n - 4 * n4;
  if (rest > 0)
    gd4d::fpe_combine_fwd_tail<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, st>>>(feat, pe, gate, sine, out,
                                                                                        4 * n4, n);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int gd4d_fpe_combine_bwd(const float* grad_out, const float* pe, const float* gate, float* grad_pe,
                         float* grad_gate, int64_t n, void* cuda_stream) {
  if (grad_out == nullptr || pe == nullptr || gate == nullptr) return GD4D_ERR_NULL;
  if (n <= 0) return GD4D_ERR_DIMS;
  gd4d::fpe_combine_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                                 static_cast<cudaStream_t>(cuda_stream)>>>(grad_out, pe, gate, grad_pe, grad_gate, n);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

}  // extern "C"
