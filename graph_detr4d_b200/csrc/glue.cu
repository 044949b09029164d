// Fused per-layer glue kernels (include/gd4d_glue.h): each replaces a chain of 5-20
// launch-latency-bound one-line torch ops of the reference's decoder layer with ONE launch.
// All fp32; the arithmetic follows the reference op by op (IEEE division, logf, expf; no
// fast-math), so results agree with torch to rounding.
//
//   inverse_sigmoid fwd/bwd   detr3d_transformer.py:28-43, deform3d_cross_attn.py:16-31
//   ref_update                detr3d_transformer.py:201-214
//   add_layernorm fwd/bwd     post-norm residual sums + position_encoder's Linear-LN-ReLU
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gd4d_glue.h"

namespace gd4d {

__device__ __forceinline__ float inv_sigmoid(float x, float eps, bool clamp_max) {
  const float xc = fminf(fmaxf(x, 0.f), 1.f);
  float x1 = fmaxf(xc, eps);
  float x2 = fmaxf(1.f - xc, eps);
  if (clamp_max) { x1 = fminf(x1, 1.f); x2 = fminf(x2, 1.f); }
  return logf(__fdiv_rn(x1, x2));
}

__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

__global__ void inverse_sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                           int64_t n, float eps, bool clamp_max) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = inv_sigmoid(x[i], eps, clamp_max);
}

__global__ void inverse_sigmoid_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                           float* __restrict__ gx, int64_t n, float eps,
                                           bool clamp_max) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i];
  const float xc = fminf(fmaxf(xv, 0.f), 1.f);
  const float om = 1.f - xc;
  // clamp(min[,max]) passes the gradient where min <= v (<= max); xc, 1-xc <= 1 always
  const float x1 = fmaxf(xc, eps), x2 = fmaxf(om, eps);
  const float m0 = (xv >= 0.f && xv <= 1.f) ? 1.f : 0.f;
  const float t1 = (xc >= eps) ? __fdiv_rn(1.f, x1) : 0.f;
  const float t2 = (om >= eps) ? __fdiv_rn(1.f, x2) : 0.f;
  (void)clamp_max;
  gx[i] = gy[i] * m0 * (t1 + t2);
}

__global__ void ref_update_kernel(const float* __restrict__ reg, int reg_stride,
                                  const float* __restrict__ ref, float* __restrict__ out,
                                  int64_t rows, float eps) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * 3) return;
  const int64_t r = i / 3;
  const int c = static_cast<int>(i - r * 3);
  const float t = reg[r * reg_stride + (c == 2 ? 4 : c)];
  out[i] = sigmoid_ref(t + inv_sigmoid(ref[i], eps, false));
}

// y = [relu](y + bias) in place, float4 per thread (C % 4 == 0)
__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, int64_t n4,
                                int C, bool relu) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = static_cast<int>((i * 4) % C);
  float4 v = reinterpret_cast<float4*>(y)[i];
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  reinterpret_cast<float4*>(y)[i] = v;
}

// ---------------------------------------------------------------------------------------
// (residual sum ->) LayerNorm (-> ReLU): one warp per row, NV float4 per lane (C = 128*NV)
// ---------------------------------------------------------------------------------------
constexpr int kLnWarps = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xbias,
                         const float* __restrict__ r1, const float* __restrict__ r2,
                         const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ pos,
                         float* __restrict__ y, float* __restrict__ y2,
                         float* __restrict__ yc1, float* __restrict__ yc2,
                         float* __restrict__ s_out, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out, int64_t rows, float eps, bool relu) {
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = static_cast<size_t>(row) * C;
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    v[j] = *reinterpret_cast<const float4*>(x + base + c);
    if (xbias != nullptr) {  // bias of the Linear that produced x (its GEMM ran without epilogue)
      const float4 a = __ldg(reinterpret_cast<const float4*>(xbias + c));
      v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
    }
    if (r1 != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(r1 + base + c);
      v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
    }
    if (r2 != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(r2 + base + c);
      v[j].x += a.x; v[j].y += a.y; v[j].z += a.z; v[j].w += a.w;
    }
    if (s_out != nullptr) *reinterpret_cast<float4*>(s_out + base + c) = v[j];
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(sum) * (1.f / C);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = __fdiv_rn(1.f, sqrtf(warp_sum(sq) * (1.f / C) + eps));
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 o;
    o.x = (v[j].x - mean) * rstd * g.x + b.x;
    o.y = (v[j].y - mean) * rstd * g.y + b.y;
    o.z = (v[j].z - mean) * rstd * g.z + b.z;
    o.w = (v[j].w - mean) * rstd * g.w + b.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4*>(y + base + c) = o;
    // identical copies of y for further consumers: each autograd consumer then hands back its own gradient and
    // the backward kernel sums them in registers, instead of one elementwise add launch per extra consumer
    if (yc1 != nullptr) *reinterpret_cast<float4*>(yc1 + base + c) = o;
    if (yc2 != nullptr) *reinterpret_cast<float4*>(yc2 + base + c) = o;
    if (y2 != nullptr) {  // second output y + pos: the next block's "query + query_pos"
      const float4 q = *reinterpret_cast<const float4*>(pos + base + c);
      o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
      *reinterpret_cast<float4*>(y2 + base + c) = o;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ gy2,
                         const float* __restrict__ gc1, const float* __restrict__ gc2,
                         const float* __restrict__ s,
                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ gs, float* __restrict__ g_masked, int64_t rows,
                         bool relu) {
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = static_cast<size_t>(row) * C;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float xh[NV][4], a[NV][4];
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 sv = *reinterpret_cast<const float4*>(s + base + c);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy != nullptr) g = *reinterpret_cast<const float4*>(gy + base + c);
    if (gy2 != nullptr) {  // gradient of the second output (y + pos)
      const float4 h = *reinterpret_cast<const float4*>(gy2 + base + c);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
    }
    if (gc1 != nullptr) {  // gradients of the copies of y
      const float4 h = *reinterpret_cast<const float4*>(gc1 + base + c);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
    }
    if (gc2 != nullptr) {
      const float4 h = *reinterpret_cast<const float4*>(gc2 + base + c);
      g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
    }
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    xh[j][0] = (sv.x - mean) * rstd; xh[j][1] = (sv.y - mean) * rstd;
    xh[j][2] = (sv.z - mean) * rstd; xh[j][3] = (sv.w - mean) * rstd;
    if (relu) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      // same expression as the forward, so the mask is the forward's [y > 0] exactly
      if (!((sv.x - mean) * rstd * gm.x + b.x > 0.f)) g.x = 0.f;
      if (!((sv.y - mean) * rstd * gm.y + b.y > 0.f)) g.y = 0.f;
      if (!((sv.z - mean) * rstd * gm.z + b.z > 0.f)) g.z = 0.f;
      if (!((sv.w - mean) * rstd * gm.w + b.w > 0.f)) g.w = 0.f;
    }
    // the EFFECTIVE incoming gradient (masked and/or summed), for the deferred gamma/beta reduction
    if (g_masked != nullptr) *reinterpret_cast<float4*>(g_masked + base + c) = g;
    a[j][0] = g.x * gm.x; a[j][1] = g.y * gm.y; a[j][2] = g.z * gm.z; a[j][3] = g.w * gm.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c1 += a[j][i]; c2 += a[j][i] * xh[j][i]; }
  }
  c1 = warp_sum(c1) * (1.f / C);
  c2 = warp_sum(c2) * (1.f / C);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    float4 o;
    o.x = rstd * (a[j][0] - c1 - xh[j][0] * c2);
    o.y = rstd * (a[j][1] - c1 - xh[j][1] * c2);
    o.z = rstd * (a[j][2] - c1 - xh[j][2] * c2);
    o.w = rstd * (a[j][3] - c1 - xh[j][3] * c2);
    *reinterpret_cast<float4*>(gs + base + c) = o;
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int launched() { return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA; }

// Softmax backward over the last dim, one warp per row, ONE pass: ds = p * (g - sum_j g_j p_j).
// (ATen materialises g * p first -- a 26 MB elementwise pass per decoder layer at 8 x 900 x 900 -- and
// then runs its warp kernel over that and p.)  Up to 1024 columns; in place over g allowed.
template <int VPL>   // float4 per lane
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ p,
                                                          float* __restrict__ ds, int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* g4 = reinterpret_cast<const float4*>(g + row * cols);
  const float4* p4 = reinterpret_cast<const float4*>(p + row * cols);
  float4* d4 = reinterpret_cast<float4*>(ds + row * cols);
  const int n4 = cols >> 2;
  float4 gv[VPL], pv[VPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int j = i * 32 + lane;
    gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    pv[i] = gv[i];
    if (j < n4) { gv[i] = g4[j]; pv[i] = __ldg(p4 + j); }
    sum += gv[i].x * pv[i].x + gv[i].y * pv[i].y + gv[i].z * pv[i].z + gv[i].w * pv[i].w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int j = i * 32 + lane;
    if (j < n4)
      d4[j] = make_float4(pv[i].x * (gv[i].x - sum), pv[i].y * (gv[i].y - sum), pv[i].z * (gv[i].z - sum),
                          pv[i].w * (gv[i].w - sum));
  }
}

}  // namespace gd4d

extern "C" {

int gd4d_inverse_sigmoid_fwd(const float* x, float* y, int64_t n, float eps, int32_t clamp_max,
                             void* cuda_stream) {
  if (x == nullptr || y == nullptr) return GD4D_ERR_NULL;
  if (n <= 0 || n > (1LL << 40)) return GD4D_ERR_DIMS;
  gd4d::inverse_sigmoid_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                                     static_cast<cudaStream_t>(cuda_stream)>>>(x, y, n, eps, clamp_max != 0);
  return gd4d::launched();
}

int gd4d_inverse_sigmoid_bwd(const float* x, const float* gy, float* gx, int64_t n, float eps,
                             int32_t clamp_max, void* cuda_stream) {
  if (x == nullptr || gy == nullptr || gx == nullptr) return GD4D_ERR_NULL;
  if (n <= 0 || n > (1LL << 40)) return GD4D_ERR_DIMS;
  gd4d::inverse_sigmoid_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                                     static_cast<cudaStream_t>(cuda_stream)>>>(x, gy, gx, n, eps,
                                                                               clamp_max != 0);
  return gd4d::launched();
}

int gd4d_ref_update(const float* reg, int32_t reg_stride, const float* ref, float* new_ref,
                    int64_t rows, float eps, void* cuda_stream) {
  if (reg == nullptr || ref == nullptr || new_ref == nullptr) return GD4D_ERR_NULL;
  if (rows <= 0 || rows > (1LL << 38) || reg_stride < 5) return GD4D_ERR_DIMS;
  const int64_t n = rows * 3;
  gd4d::ref_update_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                            static_cast<cudaStream_t>(cuda_stream)>>>(reg, reg_stride, ref, new_ref, rows, eps);
  return gd4d::launched();
}

int gd4d_bias_act(float* y, const float* bias, int64_t rows, int32_t C, int32_t relu,
                  void* cuda_stream) {
  if (y == nullptr || bias == nullptr) return GD4D_ERR_NULL;
  if (rows <= 0 || C <= 0 || C % 4 != 0 || rows > (1LL << 40) / C) return GD4D_ERR_DIMS;
  if (!gd4d::al16(y) || !gd4d::al16(bias)) return GD4D_ERR_ALIGN;
  const int64_t n4 = rows * C / 4;
  gd4d::bias_act_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(cuda_stream)>>>(y, bias, n4, C, relu != 0);
  return gd4d::launched();
}

int gd4d_add_layernorm_fwd(const float* x, const float* xbias, const float* r1, const float* r2,
                           const float* gamma, const float* beta, const float* pos, float* y,
                           float* y2, float* y_copy1, float* y_copy2, float* s_out, float* mean, float* rstd,
                           int64_t rows, int32_t C, float eps, int32_t relu, void* cuda_stream) {
  if (!gd4d::al16(y_copy1) || !gd4d::al16(y_copy2)) return GD4D_ERR_ALIGN;
  if ((pos == nullptr) != (y2 == nullptr)) return GD4D_ERR_NULL;
  if (!gd4d::al16(pos) || !gd4d::al16(y2)) return GD4D_ERR_ALIGN;
  if (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr || mean == nullptr ||
      rstd == nullptr)
    return GD4D_ERR_NULL;
  if (rows <= 0 || rows > (1LL << 31) || C <= 0 || C % 128 != 0 || C > 1024) return GD4D_ERR_DIMS;
  if ((xbias != nullptr || r1 != nullptr || r2 != nullptr) && s_out == nullptr)
    return GD4D_ERR_NULL;  // the backward needs s whenever it differs from x
  if (!gd4d::al16(x) || !gd4d::al16(xbias) || !gd4d::al16(r1) || !gd4d::al16(r2) ||
      !gd4d::al16(gamma) || !gd4d::al16(beta) || !gd4d::al16(y) || !gd4d::al16(s_out))
    return GD4D_ERR_ALIGN;
  const unsigned grid = static_cast<unsigned>((rows + gd4d::kLnWarps - 1) / gd4d::kLnWarps);
  const int block = gd4d::kLnWarps * 32;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
#define GD4D_LN_FWD(NV)                                                                          \
  gd4d::add_layernorm_fwd_kernel<NV><<<grid, block, 0, st>>>(x, xbias, r1, r2, gamma, beta, pos, y, \
                                                             y2, y_copy1, y_copy2, s_out, mean, rstd, \
                                                             rows, eps, relu != 0)
  switch (C / 128) {
    case 1: GD4D_LN_FWD(1); break;
    case 2: GD4D_LN_FWD(2); break;
    case 3: GD4D_LN_FWD(3); break;
    case 4: GD4D_LN_FWD(4); break;
    case 5: GD4D_LN_FWD(5); break;
    case 6: GD4D_LN_FWD(6); break;
    case 7: GD4D_LN_FWD(7); break;
    default: GD4D_LN_FWD(8); break;
  }
#undef GD4D_LN_FWD
  return gd4d::launched();
}

int gd4d_add_layernorm_bwd(const float* gy, const float* gy2, const float* g_copy1, const float* g_copy2,
                           const float* s, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, float* gs, float* g_masked,
                           int64_t rows, int32_t C, int32_t relu, void* cuda_stream) {
  if ((gy == nullptr && gy2 == nullptr && g_copy1 == nullptr && g_copy2 == nullptr) || s == nullptr ||
      mean == nullptr || rstd == nullptr || gamma == nullptr || gs == nullptr || (relu && beta == nullptr))
    return GD4D_ERR_NULL;
  if (!gd4d::al16(g_copy1) || !gd4d::al16(g_copy2)) return GD4D_ERR_ALIGN;
  if (rows <= 0 || rows > (1LL << 31) || C <= 0 || C % 128 != 0 || C > 1024) return GD4D_ERR_DIMS;
  if (!gd4d::al16(gy) || !gd4d::al16(gy2) || !gd4d::al16(s) || !gd4d::al16(gamma) || !gd4d::al16(beta) ||
      !gd4d::al16(gs) || !gd4d::al16(g_masked))
    return GD4D_ERR_ALIGN;
  const unsigned grid = static_cast<unsigned>((rows + gd4d::kLnWarps - 1) / gd4d::kLnWarps);
  const int block = gd4d::kLnWarps * 32;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
#define GD4D_LN_BWD(NV)                                                                          \
  gd4d::add_layernorm_bwd_kernel<NV><<<grid, block, 0, st>>>(gy, gy2, g_copy1, g_copy2, s, mean, rstd, gamma, \
                                                             beta, gs, g_masked, rows, relu != 0)
  switch (C / 128) {
    case 1: GD4D_LN_BWD(1); break;
    case 2: GD4D_LN_BWD(2); break;
    case 3: GD4D_LN_BWD(3); break;
    case 4: GD4D_LN_BWD(4); break;
    case 5: GD4D_LN_BWD(5); break;
    case 6: GD4D_LN_BWD(6); break;
    case 7: GD4D_LN_BWD(7); break;
    default: GD4D_LN_BWD(8); break;
  }
#undef GD4D_LN_BWD
  return gd4d::launched();
}

int gd4d_softmax_bwd(const float* grad_out, const float* probs, float* grad_in, int64_t rows, int32_t cols,
                     void* cuda_stream) {
  if (grad_out == nullptr || probs == nullptr || grad_in == nullptr) return GD4D_ERR_NULL;
  if (rows <= 0 || cols <= 0 || cols % 4 != 0 || cols > 1024 || rows > (1LL << 34)) return GD4D_ERR_DIMS;
  if (!gd4d::al16(grad_out) || !gd4d::al16(probs) || !gd4d::al16(grad_in)) return GD4D_ERR_ALIGN;
  auto st = static_cast<cudaStream_t>(cuda_stream);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  const int n4 = cols / 4;
  if (n4 <= 64) gd4d::softmax_bwd_kernel<2><<<grid, 256, 0, st>>>(grad_out, probs, grad_in, rows, cols);
  else if (n4 <= 128) gd4d::softmax_bwd_kernel<4><<<grid, 256, 0, st>>>(grad_out, probs, grad_in, rows, cols);
  else gd4d::softmax_bwd_kernel<8><<<grid, 256, 0, st>>>(grad_out, probs, grad_in, rows, cols);
  return gd4d::launched();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// AdamW over MANY small tensors in ONE launch (include/gd4d_glue.h: gd4d_adamw_multi).
// torch's fused AdamW walks 65536-element chunks, one CTA each: the decoder's ~7 M parameters in
// ~200 tensors become ~150 CTAs spread over 6 launches (39 us each, r1).  Here a CTA owns a
// 4096-element chunk (block map built once on the host), so the same update is ~2000 CTAs in a
// single launch, HBM-bound (28 B per element).
// Math = torch.optim.AdamW(fused=True): decoupled weight decay, lerp for exp_avg, bias
// corrections from the DEVICE step counter (capturable), evaluated in double like torch.
// ---------------------------------------------------------------------------------------
namespace gd4d {
constexpr int kAdamChunk = 4096;

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const gd4d_adamw_tensor* __restrict__ table, const int2* __restrict__ block_map,
                   const float* __restrict__ step_ptr, float lr, float beta1, float beta2, float eps,
                   float weight_decay) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double step = static_cast<double>(*step_ptr);
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    s_step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    s_bc2_sqrt = static_cast<float>(sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const int2 bm = block_map[blockIdx.x];
  const gd4d_adamw_tensor t = table[bm.x];
  const int64_t start = static_cast<int64_t>(bm.y) * kAdamChunk;
  const int64_t end = min(start + kAdamChunk, t.n);
  const float decay = 1.f - lr * weight_decay;
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                     reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15u) == 0;
  auto upd = [&](float& p, float g, float& m, float& v) {
    p *= decay;
    m = m + omb1 * (g - m);                      // lerp(m, g, 1 - beta1)
    v = beta2 * v + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step_size * m / denom;
  };
  if (vec) {
    for (int64_t i = start + threadIdx.x * 4; i < end; i += 256 * 4) {
      if (i + 4 <= end) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
      } else {
        for (int64_t j = i; j < end; ++j) upd(t.p[j], t.g[j], t.m[j], t.v[j]);
      }
    }
  } else {
    for (int64_t i = start + threadIdx.x; i < end; i += 256) upd(t.p[i], t.g[i], t.m[i], t.v[i]);
  }
}
}  // namespace gd4d

extern "C" int gd4d_adamw_chunk(void) { return gd4d::kAdamChunk; }

extern "C" int gd4d_adamw_multi(const gd4d_adamw_tensor* table_dev, const int32_t* block_map_dev,
                                int32_t n_blocks, const float* step_dev, float lr, float beta1,
                                float beta2, float eps, float weight_decay, void* cuda_stream) {
  if (table_dev == nullptr || block_map_dev == nullptr || step_dev == nullptr) return GD4D_ERR_NULL;
  if (n_blocks <= 0) return GD4D_ERR_DIMS;
  gd4d::adamw_multi_kernel<<<static_cast<unsigned>(n_blocks), 256, 0,
                             static_cast<cudaStream_t>(cuda_stream)>>>(
      table_dev, reinterpret_cast<const int2*>(block_map_dev), step_dev, lr, beta1, beta2, eps,
      weight_decay);
  return gd4d::launched();
}
