// Backward of the fused cross-view sampling kernel.
//
// Same decomposition as the forward (one warp per (b, q, head); lane groups own
// (valid candidate, level) items).  Nothing from the forward is saved: the
// projection, mask and softmax are recomputed in registers (cheap), the four
// corner slices are re-gathered, and per item the group forms three channel
// dot-products with grad_out -- (s.g), (ds/dix.g), (ds/diy.g) -- by shuffle
// reduction.  From those:
//   grad_value        128-byte (fp32 grads) vector reductions `red.global.add.v4.f32`
//                     straight into the channel-last grad map: one warp-wide
//                     instruction retires 4 full corner slices
//   grad_attn_logits  softmax backward per head (C) / sigmoid' (A)
//   grad_cam_logits   sigmoid' * sum_heads(partial_out . g)            (C)
//   grad_offsets/ref  chain through u=(cx/den)/W_img ... lidar2img^T   
//
// Reference being replaced: autograd through detr3d_transformer.py:376-438 /
// deform3d_cross_attn.py:211-324, i.e. aten grid_sampler_2d_backward and mmcv
// ms_deform_attn_backward (ms_deformable_col2im_gpu_kernel_*).
// All gradient outputs ACCUMULATE into caller-zeroed buffers.
#include "xview_common.cuh"

namespace gd4d {

struct __align__(16) CandB {
  float u, v, den, w;   // w = sigmoid(cam logit) (C) or 1 (A)
  int np;               // n<<8 | p
  float du, dv, cg;     // accumulators: dL/du, dL/dv, sum_l sm[l,p]*(s.g)
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a),
               "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int VEC>
__device__ __forceinline__ void red_slice(float* addr, float w, const float (&g)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i += 4) red_add_v4(addr + i, w * g[i], w * g[i + 1], w * g[i + 2], w * g[i + 3]);
}

template <int MODE, typename VT>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
xview_bwd_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int LANES = Slice<VT>::LANES;
  constexpr int GROUPS = 32 / LANES;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = lane / LANES;
  const int sub = lane % LANES;
  const size_t warp_bytes = sizeof(float) * kMaxLP * 5 + sizeof(CandB) * cand_cap;
  float* sw = reinterpret_cast<float*>(smem_raw + warp * warp_bytes);  // softmax weights
  float* gsum = sw + kMaxLP;                                           // sum_n wcam*(s.g) per (l,p)
  float* doff = gsum + kMaxLP;                                         // dL/d offset (p,3)
  CandB* cands = reinterpret_cast<CandB*>(doff + 3 * kMaxLP);

  const long long gw = static_cast<long long>(blockIdx.x) * kWarpsPerCta + warp;
  const long long total_warps = static_cast<long long>(p.B) * p.Q * p.Hh;
  if (gw >= total_warps) return;
  const int h = static_cast<int>(gw % p.Hh);
  const int bq = static_cast<int>(gw / p.Hh);
  const int b = bq / p.Q;
  const int q = bq - b * p.Q;
  const int LP = p.L * p.P;

  const float* rp = p.ref + static_cast<size_t>(bq) * 3;
  const float X0 = __fadd_rn(__fmul_rn(__ldg(rp + 0), p.pc_span[0]), p.pc_lo[0]);
  const float Y0 = __fadd_rn(__fmul_rn(__ldg(rp + 1), p.pc_span[1]), p.pc_lo[1]);
  const float Z0 = __fadd_rn(__fmul_rn(__ldg(rp + 2), p.pc_span[2]), p.pc_lo[2]);

  // grad_out slice owned by this lane (identical in every lane group)
  float g[VEC];
  {
    const float* go = p.grad_out + static_cast<size_t>(bq) * p.C + h * kHeadDim + sub * VEC;
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(go + i));
      g[i] = t.x; g[i + 1] = t.y; g[i + 2] = t.z; g[i + 3] = t.w;
    }
  }

  if (MODE == GD4D_MODE_C) {
    const float* a = p.attn_logits + (static_cast<size_t>(bq) * p.Hh + h) * LP;
    const float x0 = lane < LP ? __ldg(a + lane) : -INFINITY;
    const float x1 = lane + 32 < LP ? __ldg(a + lane + 32) : -INFINITY;
    float m = fmaxf(x0, x1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e0 = lane < LP ? expf(x0 - m) : 0.f;
    const float e1 = lane + 32 < LP ? expf(x1 - m) : 0.f;
    float s = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    sw[lane] = e0 / s;
    sw[lane + 32] = e1 / s;
    gsum[lane] = 0.f;
    gsum[lane + 32] = 0.f;
    for (int i = lane; i < 3 * kMaxLP; i += 32) doff[i] = 0.f;
  }

  // ---- phase 1: candidates ---------------------------------------------------------
  const int PP = (MODE == GD4D_MODE_C) ? p.P : 1;
  const int ncand = p.N * PP;
  int nvalid = 0;
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    bool valid = false;
    CandB cd;
    cd.u = cd.v = 0.f; cd.den = 1.f; cd.w = 1.f; cd.np = 0; cd.du = cd.dv = cd.cg = 0.f;
    if (c < ncand) {
      const int n = c / PP;
      const int pi = c - n * PP;
      float X = X0, Y = Y0, Z = Z0;
      if (MODE == GD4D_MODE_C) {
        const float* o = p.offsets + ((static_cast<size_t>(bq) * p.Hh + h) * p.P + pi) * 3;
        X = __fadd_rn(X, __ldg(o + 0));
        Y = __fadd_rn(Y, __ldg(o + 1));
        Z = __fadd_rn(Z, __ldg(o + 2));
      }
      const float* M = p.lidar2img + (static_cast<size_t>(b) * p.N + n) * 16;
      const Projected pr = project_point(M, X, Y, Z, p.img_w, p.img_h);
      valid = pr.depth_ok & in_image<MODE>(pr.u, pr.v);
      cd.u = pr.u; cd.v = pr.v; cd.den = pr.den; cd.np = (n << 8) | pi;
      if (valid && MODE == GD4D_MODE_C)
        cd.w = sigmoidf_(__ldg(p.cam_logits + static_cast<size_t>(b) * p.N * p.Q +
                               static_cast<size_t>(n) * p.Q + q));
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (valid) cands[nvalid + __popc(bal & ((1u << lane) - 1u))] = cd;
    nvalid += __popc(bal);
  }
  __syncwarp();

  // ---- phase 2: re-gather, dot with grad_out, scatter feature gradients ---------------------
  const int total = nvalid * p.L;
  for (int it0 = 0; it0 < total; it0 += GROUPS) {
    const int item = it0 + grp;
    const bool active = item < total;
    const int it = active ? item : 0;
    const int k = it / p.L;
    const int l = it - k * p.L;
    const int np = cands[k].np;
    const float cu = cands[k].u, cv = cands[k].v, cw = cands[k].w;
    const int n = np >> 8;
    const int pi = np & 0xff;
    float wt, smw = 0.f;
    const float* alog = nullptr;
    if (MODE == GD4D_MODE_C) {
      smw = sw[l * p.P + pi];
      wt = smw * cw;
    } else {
      alog = p.attn_logits + ((static_cast<size_t>(b) * p.Q + q) * p.N + n) * p.P * p.L + l;
      wt = 0.f;
      for (int pp = 0; pp < p.P; ++pp) wt += sigmoidf_(__ldg(alog + pp * p.L));
    }
    if (!active) wt = 0.f;
    const int W = p.level_w[l], H = p.level_h[l];
    const float ix = to_pixel(to_grid<MODE>(cu), static_cast<float>(W));
    const float iy = to_pixel(to_grid<MODE>(cv), static_cast<float>(H));
    const Footprint f = footprint(ix, iy, W, H);
    const VT* base = static_cast<const VT*>(p.value[l]);
    const size_t img = static_cast<size_t>(b) * p.N + n;
    const size_t e00 = ((img * H + f.y0) * W + f.x0) * p.C + static_cast<size_t>(h) * kHeadDim + sub * VEC;
    const size_t rowst = static_cast<size_t>(W) * p.C;
    float c00[VEC], c01[VEC], c10[VEC], c11[VEC];
    const bool a00 = active & f.in00, a01 = active & f.in01, a10 = active & f.in10, a11 = active & f.in11;
    Slice<VT>::load(base + e00, a00, c00);
    Slice<VT>::load(base + e00 + p.C, a01, c01);
    Slice<VT>::load(base + e00 + rowst, a10, c10);
    Slice<VT>::load(base + e00 + rowst + p.C, a11, c11);
    const float w00 = (1.f - f.tx) * (1.f - f.ty), w01 = f.tx * (1.f - f.ty);
    const float w10 = (1.f - f.tx) * f.ty, w11 = f.tx * f.ty;

    // feature-map gradient: dL/df_c += wt * w_c * g   (vector reductions, no return value)
    float* gv = p.grad_value[l];
    if (gv != nullptr && wt != 0.f) {
      if (a00) red_slice<VEC>(gv + e00, wt * w00, g);
      if (a01) red_slice<VEC>(gv + e00 + p.C, wt * w01, g);
      if (a10) red_slice<VEC>(gv + e00 + rowst, wt * w10, g);
      if (a11) red_slice<VEC>(gv + e00 + rowst + p.C, wt * w11, g);
    }

    float sdot = 0.f, dxdot = 0.f, dydot = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sdot += g[i] * (w00 * c00[i] + w01 * c01[i] + w10 * c10[i] + w11 * c11[i]);
      dxdot += g[i] * ((c01[i] - c00[i]) * (1.f - f.ty) + (c11[i] - c10[i]) * f.ty);
      dydot += g[i] * ((c10[i] - c00[i]) * (1.f - f.tx) + (c11[i] - c01[i]) * f.tx);
    }
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) {
      sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
      dxdot += __shfl_xor_sync(0xffffffffu, dxdot, o);
      dydot += __shfl_xor_sync(0xffffffffu, dydot, o);
    }
    if (active && sub == 0) {
      atomicAdd(&cands[k].du, wt * static_cast<float>(W) * dxdot);
      atomicAdd(&cands[k].dv, wt * static_cast<float>(H) * dydot);
      if (MODE == GD4D_MODE_C) {
        atomicAdd(&gsum[l * p.P + pi], cw * sdot);
        atomicAdd(&cands[k].cg, smw * sdot);
      } else if (p.grad_attn_logits != nullptr) {
        float* ga = p.grad_attn_logits + (alog - p.attn_logits);
        for (int pp = 0; pp < p.P; ++pp) {
          const float sg = sigmoidf_(__ldg(alog + pp * p.L));
          atomicAdd(ga + pp * p.L, sg * (1.f - sg) * sdot);
        }
      }
    }
  }
  __syncwarp();

  // ---- phase 3: small gradients -----------------------------------------------------
  if (MODE == GD4D_MODE_C && p.grad_attn_logits != nullptr) {
    // softmax backward: dlogit_j = sm_j * (G_j - sum_k sm_k G_k)
    const float s0 = sw[lane], s1 = sw[lane + 32];
    const float g0 = gsum[lane], g1 = gsum[lane + 32];
    float dot = s0 * g0 + s1 * g1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    float* ga = p.grad_attn_logits + (static_cast<size_t>(bq) * p.Hh + h) * LP;
    if (lane < LP) atomicAdd(ga + lane, s0 * (g0 - dot));
    if (lane + 32 < LP) atomicAdd(ga + lane + 32, s1 * (g1 - dot));
  }

  float rX = 0.f, rY = 0.f, rZ = 0.f;
  for (int k = lane; k < nvalid; k += 32) {
    const CandB cd = cands[k];
    const int n = cd.np >> 8;
    const int pi = cd.np & 0xff;
    const float* M = p.lidar2img + (static_cast<size_t>(b) * p.N + n) * 16;
    const float dcx = cd.du / (cd.den * p.img_w);
    const float dcy = cd.dv / (cd.den * p.img_h);
    const float dcz = -(cd.du * cd.u + cd.dv * cd.v) / cd.den;  // valid => cz > eps => d den/d cz = 1
    const float dX = __ldg(M + 0) * dcx + __ldg(M + 4) * dcy + __ldg(M + 8) * dcz;
    const float dY = __ldg(M + 1) * dcx + __ldg(M + 5) * dcy + __ldg(M + 9) * dcz;
    const float dZ = __ldg(M + 2) * dcx + __ldg(M + 6) * dcy + __ldg(M + 10) * dcz;
    rX += dX; rY += dY; rZ += dZ;
    if (MODE == GD4D_MODE_C) {
      atomicAdd(&doff[pi * 3 + 0], dX);
      atomicAdd(&doff[pi * 3 + 1], dY);
      atomicAdd(&doff[pi * 3 + 2], dZ);
      if (p.grad_cam_logits != nullptr)
        atomicAdd(p.grad_cam_logits + static_cast<size_t>(b) * p.N * p.Q + static_cast<size_t>(n) * p.Q + q,
                  cd.w * (1.f - cd.w) * cd.cg);
    }
  }
  if (p.grad_ref != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rX += __shfl_xor_sync(0xffffffffu, rX, o);
      rY += __shfl_xor_sync(0xffffffffu, rY, o);
      rZ += __shfl_xor_sync(0xffffffffu, rZ, o);
    }
    if (lane == 0 && nvalid > 0) {
      float* gr = p.grad_ref + static_cast<size_t>(bq) * 3;
      atomicAdd(gr + 0, rX * p.pc_span[0]);
      atomicAdd(gr + 1, rY * p.pc_span[1]);
      atomicAdd(gr + 2, rZ * p.pc_span[2]);
    }
  }
  if (MODE == GD4D_MODE_C && p.grad_offsets != nullptr) {
    __syncwarp();
    float* go = p.grad_offsets + (static_cast<size_t>(bq) * p.Hh + h) * p.P * 3;
    for (int i = lane; i < p.P * 3; i += 32) atomicAdd(go + i, doff[i]);
  }
}

template <int MODE, typename VT>
static int launch_bwd(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_bwd_kernel<MODE, VT>;
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem);
    if (e != cudaSuccess) return GD4D_ERR_CUDA;
  }
  kern<<<g.grid, g.block, g.smem, stream>>>(p, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int dispatch_backward(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  const bool bf16 = p.value_dtype == GD4D_BF16;
  switch (p.mode) {
    case GD4D_MODE_A:
      return bf16 ? launch_bwd<GD4D_MODE_A, __nv_bfloat16>(p, g, stream)
                  : launch_bwd<GD4D_MODE_A, float>(p, g, stream);
    case GD4D_MODE_C:
      return bf16 ? launch_bwd<GD4D_MODE_C, __nv_bfloat16>(p, g, stream)
                  : launch_bwd<GD4D_MODE_C, float>(p, g, stream);
    default:
      return GD4D_ERR_UNSUPPORTED;
  }
}

}  // namespace gd4d
