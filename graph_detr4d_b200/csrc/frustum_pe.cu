// Fused frustum position-embedding input (include/gd4d_frustum.h).
// Replaces the elementwise body of Detr3DHeadPE.position_embeding
// (dense_heads/detr3d_head_pe.py:439-480): one thread per (camera image, pixel) walks the D depth
// bins, lifts the frustum point through img2lidar with explicit round-to-nearest mul/add in the
// reference's order (so the out-of-range count, hence the mask, is bit-exact), normalises,
// applies inverse_sigmoid and writes the (B*N, D*3, H, W) convolution input directly:
// consecutive threads own consecutive pixels, so every store instruction of a warp is one
// contiguous 128-byte line of one (d, c) plane.  HBM-write-bound: 12 B per (pixel, depth bin).
//
// r2: the r1 kernel spent ~45 instructions per coordinate (two IEEE divisions with their slow-path
// checks and an accurate logf) and ran at 0.17 of the HBM copy peak, issue-bound.  The MASK needs the
// reference's exact arithmetic, the VALUE does not (the reference's own BLAS-ordered mat-vec already
// moves it by ~1e-6 in the normalised domain, tests/test_pe_oracle.py): so
//   * the out-of-range test is decided on t = acc - lo without dividing: for span > 0,
//     fl(t / span) > 1  <=>  t > span   (t = nextafter(span) divides to 1 + 2^-23/m, m in [1,2): rounds up)
//     fl(t / span) < 0  <=>  t < 0      (except when the quotient underflows to -0: that rare case takes
//                                        the exact division)
//     and when the camera matrix is bounded and |lo| is a normal-sized number (every real pc_range), t is
//     never -0 / nan / denormal-tiny, so both tests are ONE unsigned compare bits(t) > bits(span)
//   * the value needs c = fl(t / span) itself: the logit amplifies one ulp of c by 1/(1-c) (up to 1e5 at
//     the clamp), so t * (1/span) alone is NOT enough (measured 1.8e-3 off).  With r = RN(1/span) from the
//     host, q = RN(t*r), e = fma(-q, span, t) (exact), c = RN(q + e*r) is the correctly rounded quotient
//     (Markstein; holds unless span's significand is all ones, which the host entry refuses) -- 3 FP32
//     ops instead of the ~10-instruction IEEE division with its slow-path check.  The final
//     log(x1/x2) uses MUFU.RCP/MUFU.LG2 (relative error ~3e-7 on the ratio -> ~3e-7 absolute on the logit).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/gd4d_frustum.h"

namespace gd4d {

constexpr int kMaxLevels = 8;

struct FrustumLevel {
  const uint8_t* mask_in;
  float* out;
  uint8_t* mask_out;
  int H, W;
  int block0;          // first blockIdx.x of this level
  int pix_per_thread;  // 4 when H*W % 4 == 0 and out is 16-byte aligned, else 1
};

struct FrustumArgs {
  const float* img2lidar;
  FrustumLevel lv[kMaxLevels];
  int num_levels, D;
  float pad_h, pad_w, depth_start, bin_size;
  float lo[3], span[3], rspan[3];
  int lo_is_normal;    // every |lo| >= 1e-20: t = acc - lo is then 0 or >= ulp(lo)/2, never -0 / denormal-tiny
};

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One (pixel, depth bin, coordinate) value.  FAST is taken when the camera matrix is finite and bounded
// (no inf / nan can appear) and every |lo| is a normal-sized number, so that
//   (c > 1) | (c < 0)  <=>  t > span | t < 0  <=>  bits(t) >u bits(span)      (t is never -0, never nan)
// -- one unsigned compare -- and the inf guard of the quotient correction is dead.  Otherwise the exact
// general path decides the (rare) underflow-to--0 case with a real division.
template <bool FAST>
__device__ __forceinline__ float frustum_value(float acc, float lo, float span, float rspan, int& outside) {
  const float eps = 1e-5f;
  const float t = __fsub_rn(acc, lo);                                                                // :469-474
  const float q = __fmul_rn(t, rspan);
  float xc;
  if (FAST) {
    outside += __float_as_uint(t) > __float_as_uint(span) ? 1 : 0;                                   // :476
    xc = __saturatef(__fmaf_rn(__fmaf_rn(-q, span, t), rspan, q));                                   // clamp(fl(t / span), 0, 1)  :480
  } else {
    bool neg = t < 0.f;
    if (neg && t > -1e-30f) neg = __fdiv_rn(t, span) < 0.f;        // quotient may underflow to -0: exact path
    outside += (t > span) | neg;                                   // == (c > 1.0) | (c < 0.0), c = fl(t/span)
    float c = __fmaf_rn(__fmaf_rn(-q, span, t), rspan, q);         // == fl(t / span)
    if (!(fabsf(q) < 1e30f)) c = q;                                // inf / nan projections: no inf - inf in the correction
    xc = __saturatef(c);                                           // nan -> 0 like fmin(fmax(nan, 0), 1)
  }
  const float x1 = fmaxf(xc, eps), x2 = fmaxf(__fsub_rn(1.f, xc), eps);
  return (lg2_approx(x1) - lg2_approx(x2)) * 0.693147180559945309f;   // log(x1 / x2); x1, x2 in [1e-5, 1]
}

template <int PIX, bool FAST>
__device__ __forceinline__ void frustum_pixels(const FrustumArgs& a, const FrustumLevel& L, const float (&M)[12],
                                               int bn, int pix0) {
  const int HW = L.H * L.W;
  const float eps = 1e-5f;
  float xw[PIX], yh[PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int pix = min(pix0 + j, HW - 1);
    const int h = pix / L.W, w = pix - h * L.W;
    xw[j] = __fdiv_rn(__fmul_rn(static_cast<float>(w), a.pad_w), static_cast<float>(L.W));   // :440
    yh[j] = __fdiv_rn(__fmul_rn(static_cast<float>(h), a.pad_h), static_cast<float>(L.H));   // :439
  }
  float* o = L.out + (static_cast<size_t>(bn) * 3 * a.D) * HW + pix0;
  int outside[PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) outside[j] = 0;
  for (int d = 0; d < a.D; ++d) {
    const float idx = static_cast<float>(d);
    const float z = __fadd_rn(a.depth_start, __fmul_rn(__fmul_rn(a.bin_size, idx), __fadd_rn(idx, 1.f)));  // :455
    const float s = fmaxf(z, eps);                                                                   // :460
    float px[PIX], py[PIX];
#pragma unroll
    for (int j = 0; j < PIX; ++j) { px[j] = __fmul_rn(xw[j], s); py[j] = __fmul_rn(yh[j], s); }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float mz = __fmul_rn(M[4 * r + 2], z);                 // shared by the PIX pixels of this thread
      float v[PIX];
#pragma unroll
      for (int j = 0; j < PIX; ++j) {
        float acc = __fmul_rn(M[4 * r + 0], px[j]);                                                  // :468, the reference's order
        acc = __fadd_rn(acc, __fmul_rn(M[4 * r + 1], py[j]));
        acc = __fadd_rn(acc, mz);
        acc = __fadd_rn(acc, M[4 * r + 3]);
        v[j] = frustum_value<FAST>(acc, a.lo[r], a.span[r], a.rspan[r], outside[j]);
      }
      float* op = o + static_cast<size_t>(d * 3 + r) * HW;
      if (PIX == 4) {
        *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < PIX; ++j) op[j] = v[j];
      }
    }
  }
  if (L.mask_out != nullptr) {
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const size_t mi = static_cast<size_t>(bn) * HW + pix0 + j;
      const bool m = static_cast<float>(outside[j]) > static_cast<float>(a.D) * 0.5f;                // :477
      L.mask_out[mi] = (m || (L.mask_in != nullptr && L.mask_in[mi] != 0)) ? 1 : 0;                  // :478
    }
  }
}

// All levels in ONE launch: blockIdx.x walks the levels' pixel blocks back to back (the small levels are
// latency-bound on their own: 36 CTAs for 15x25), blockIdx.y = camera image.  PIX consecutive pixels per
// thread; with PIX == 4 every store is one 16-byte st.global.v4, i.e. a warp retires 512 contiguous bytes of
// a (d, c) plane per instruction, a 128-thread CTA 2 KB.
__global__ void __launch_bounds__(128) frustum_pe_kernel(const __grid_constant__ FrustumArgs a) {
  int l = 0;
#pragma unroll 1
  while (l + 1 < a.num_levels && static_cast<int>(blockIdx.x) >= a.lv[l + 1].block0) ++l;
  const FrustumLevel& L = a.lv[l];
  const int bn = blockIdx.y;
  const int pix0 = ((blockIdx.x - L.block0) * blockDim.x + threadIdx.x) * L.pix_per_thread;
  if (pix0 >= L.H * L.W) return;
  float M[12];
  bool bounded = a.lo_is_normal != 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    M[i] = __ldg(a.img2lidar + static_cast<size_t>(bn) * 16 + i);
    bounded = bounded && fabsf(M[i]) < 1e15f;                      // false for inf / nan too
  }
  if (L.pix_per_thread == 4) {
    if (bounded) frustum_pixels<4, true>(a, L, M, bn, pix0); else frustum_pixels<4, false>(a, L, M, bn, pix0);
  } else {
    if (bounded) frustum_pixels<1, true>(a, L, M, bn, pix0); else frustum_pixels<1, false>(a, L, M, bn, pix0);
  }
}

}  // namespace gd4d

extern "C" int gd4d_frustum_pe_levels(const float* img2lidar, const uint8_t* const* mask_in, float* const* out,
                                      uint8_t* const* mask_out, int32_t BN, int32_t num_levels,
                                      const int32_t* level_h, const int32_t* level_w, int32_t D, float pad_h,
                                      float pad_w, float depth_start, float bin_size, const float* pc_lo_span,
                                      void* cuda_stream) {
  if (img2lidar == nullptr || out == nullptr || pc_lo_span == nullptr || level_h == nullptr || level_w == nullptr)
    return GD4D_ERR_NULL;
  if (BN <= 0 || BN > 65535 || D <= 0 || D > 4096 || num_levels <= 0 || num_levels > gd4d::kMaxLevels)
    return GD4D_ERR_DIMS;
  gd4d::FrustumArgs a;
  a.img2lidar = img2lidar;
  a.num_levels = num_levels; a.D = D;
  a.pad_h = pad_h; a.pad_w = pad_w; a.depth_start = depth_start; a.bin_size = bin_size;
  a.lo_is_normal = 1;
  // pad_h * h and the depth schedule stay far below the bounds the fast path assumes (|coordinate| < 1e7)
  if (!(fabsf(pad_h) < 1e6f && fabsf(pad_w) < 1e6f && fabsf(depth_start) < 1e6f &&
        fabsf(bin_size) * static_cast<float>(D) * static_cast<float>(D + 1) < 1e6f))
    a.lo_is_normal = 0;
  for (int i = 0; i < 3; ++i) {
    a.lo[i] = pc_lo_span[i]; a.span[i] = pc_lo_span[3 + i];
    if (!(a.span[i] > 0.f)) return GD4D_ERR_DIMS;     // the division-free range test needs span > 0
    uint32_t bits;
    memcpy(&bits, &a.span[i], 4);
    if ((bits & 0x7fffffu) == 0x7fffffu) return GD4D_ERR_DIMS;   // Markstein's correction needs a non-all-ones significand
    a.rspan[i] = static_cast<float>(1.0 / static_cast<double>(a.span[i]));
    if (!(fabsf(a.lo[i]) >= 1e-20f && fabsf(a.lo[i]) < 1e15f && a.span[i] < 1e15f)) a.lo_is_normal = 0;
  }
  long long blocks = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int H = level_h[l], W = level_w[l];
    if (H <= 0 || W <= 0 || static_cast<long long>(H) * W > 0x7fffffffLL) return GD4D_ERR_DIMS;
    if (out[l] == nullptr) return GD4D_ERR_NULL;
    gd4d::FrustumLevel& L = a.lv[l];
    L.mask_in = mask_in != nullptr ? mask_in[l] : nullptr;
    L.out = out[l];
    L.mask_out = mask_out != nullptr ? mask_out[l] : nullptr;
    L.H = H; L.W = W;
    L.pix_per_thread = ((H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(out[l]) & 15u) == 0) ? 4 : 1;
    L.block0 = static_cast<int>(blocks);
    blocks += (H * W / L.pix_per_thread + 127) / 128;
    if (blocks > 0x7fffffffLL) return GD4D_ERR_DIMS;
  }
  dim3 grid(static_cast<unsigned>(blocks), BN);
  gd4d::frustum_pe_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

extern "C" int gd4d_frustum_pe(const float* img2lidar, const uint8_t* mask_in, float* out,
                               uint8_t* mask_out, int32_t BN, int32_t H, int32_t W, int32_t D,
                               float pad_h, float pad_w, float depth_start, float bin_size,
                               const float* pc_lo_span, void* cuda_stream) {
  if (img2lidar == nullptr || out == nullptr || pc_lo_span == nullptr) return GD4D_ERR_NULL;
  const uint8_t* mi[1] = {mask_in};
  float* o[1] = {out};
  uint8_t* mo[1] = {mask_out};
  const int32_t hs[1] = {H}, ws[1] = {W};
  return gd4d_frustum_pe_levels(img2lidar, mi, o, mo, BN, 1, hs, ws, D, pad_h, pad_w, depth_start, bin_size,
                                pc_lo_span, cuda_stream);
}
