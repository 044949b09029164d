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

struct FrustumArgs {
  const float* img2lidar;
  const uint8_t* mask_in;
  float* out;
  uint8_t* mask_out;
  int H, W, D;
  float pad_h, pad_w, depth_start, bin_size;
  float lo[3], span[3], rspan[3];
};

__global__ void __launch_bounds__(128) frustum_pe_kernel(const FrustumArgs a) {
  const int HW = a.H * a.W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const int bn = blockIdx.y;
  const int h = pix / a.W, w = pix - h * a.W;
  const float eps = 1e-5f;
  const float xw = __fdiv_rn(__fmul_rn(static_cast<float>(w), a.pad_w), static_cast<float>(a.W));   // :440
  const float yh = __fdiv_rn(__fmul_rn(static_cast<float>(h), a.pad_h), static_cast<float>(a.H));   // :439
  float M[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) M[i] = __ldg(a.img2lidar + static_cast<size_t>(bn) * 16 + i);
  float* o = a.out + (static_cast<size_t>(bn) * 3 * a.D) * HW + pix;
  int outside = 0;
  for (int d = 0; d < a.D; ++d) {
    const float idx = static_cast<float>(d);
    const float z = __fadd_rn(a.depth_start, __fmul_rn(__fmul_rn(a.bin_size, idx), __fadd_rn(idx, 1.f)));  // :455
    const float s = fmaxf(z, eps);                                                                   // :460
    const float px = __fmul_rn(xw, s), py = __fmul_rn(yh, s);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = __fmul_rn(M[4 * r + 0], px);                                                       // :468
      acc = __fadd_rn(acc, __fmul_rn(M[4 * r + 1], py));
      acc = __fadd_rn(acc, __fmul_rn(M[4 * r + 2], z));
      acc = __fadd_rn(acc, M[4 * r + 3]);
      const float t = __fsub_rn(acc, a.lo[r]);                                                       // :469-474
      bool neg = t < 0.f;
      if (neg && t > -1e-30f) neg = __fdiv_rn(t, a.span[r]) < 0.f;   // quotient may underflow to -0: exact path
      outside += (t > a.span[r]) | neg;                              // == (c > 1.0) | (c < 0.0), c = fl(t/span)  :476
      const float q = __fmul_rn(t, a.rspan[r]);
      float c = __fmaf_rn(__fmaf_rn(-q, a.span[r], t), a.rspan[r], q);                              // == fl(t / span)
      if (!(fabsf(q) < 1e30f)) c = q;                              // inf / nan projections: no inf - inf in the correction
      const float xc = fminf(fmaxf(c, 0.f), 1.f);                                                    // :480
      const float x1 = fmaxf(xc, eps), x2 = fmaxf(1.f - xc, eps);
      o[static_cast<size_t>(d * 3 + r) * HW] = __logf(__fdividef(x1, x2));
    }
  }
  if (a.mask_out != nullptr) {
    const size_t mi = static_cast<size_t>(bn) * HW + pix;
    const bool m = static_cast<float>(outside) > static_cast<float>(a.D) * 0.5f;                     // :477
    a.mask_out[mi] = (m || (a.mask_in != nullptr && a.mask_in[mi] != 0)) ? 1 : 0;                    // :478
  }
}

}  // namespace gd4d

extern "C" int gd4d_frustum_pe(const float* img2lidar, const uint8_t* mask_in, float* out,
                               uint8_t* mask_out, int32_t BN, int32_t H, int32_t W, int32_t D,
                               float pad_h, float pad_w, float depth_start, float bin_size,
                               const float* pc_lo_span, void* cuda_stream) {
  if (img2lidar == nullptr || out == nullptr || pc_lo_span == nullptr) return GD4D_ERR_NULL;
  if (BN <= 0 || BN > 65535 || H <= 0 || W <= 0 || D <= 0 || D > 4096) return GD4D_ERR_DIMS;
  if (static_cast<long long>(H) * W > 0x7fffffffLL) return GD4D_ERR_DIMS;
  gd4d::FrustumArgs a;
  a.img2lidar = img2lidar; a.mask_in = mask_in; a.out = out; a.mask_out = mask_out;
  a.H = H; a.W = W; a.D = D;
  a.pad_h = pad_h; a.pad_w = pad_w; a.depth_start = depth_start; a.bin_size = bin_size;
  for (int i = 0; i < 3; ++i) {
    a.lo[i] = pc_lo_span[i]; a.span[i] = pc_lo_span[3 + i];
    if (!(a.span[i] > 0.f)) return GD4D_ERR_DIMS;     // the division-free range test needs span > 0
    uint32_t bits;
    memcpy(&bits, &a.span[i], 4);
    if ((bits & 0x7fffffu) == 0x7fffffu) return GD4D_ERR_DIMS;   // Markstein's correction needs a non-all-ones significand
    a.rspan[i] = static_cast<float>(1.0 / static_cast<double>(a.span[i]));
  }
  dim3 grid((H * W + 127) / 128, BN);
  gd4d::frustum_pe_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}
