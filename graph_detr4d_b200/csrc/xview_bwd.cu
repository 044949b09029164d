// Backward of the fused cross-view sampling kernel.
//
// Same decomposition as the forward (one warp per (b, q, head); lane groups own
// (valid candidate, level) items; narrow / wide lane layouts as in xview_fwd.cu).
// Nothing from the forward is saved: the projection, mask and softmax are
// recomputed in registers (cheap), the four corner runs are re-gathered, and per
// item the group forms three channel dot-products with grad_out -- (s.g),
// (ds/dix.g), (ds/diy.g) -- by shuffle reduction.  From those:
//   grad_value        16-byte-per-lane vector reductions `red.global.add.v4.f32`
//                     straight into the channel-last fp32 grad map: one warp-wide
//                     instruction retires 512 contiguous bytes per corner run
//   grad_attn_logits  softmax backward per head (C) / sigmoid' (A)
//   grad_cam_logits   sigmoid' * sum_heads(partial_out . g)            (C)
//   grad_offsets/ref  chain through u=(cx/den)/W_img ... lidar2img^T   
// In wide mode the value-bias term rides along as one extra "all-ones" channel
// whose gradient is grad_wsum.
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
  __device__ __forceinline__ static CandB make(const Projected& pr, int n, int pi, float wc) {
    CandB c;
    c.u = pr.u; c.v = pr.v; c.den = pr.den; c.w = wc; c.np = (n << 8) | pi;
    c.du = c.dv = c.cg = 0.f;
    return c;
  }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a),
               "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int MODE, typename VT, int LANES, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (LANES == 32) ? 2 : 3)
xview_bwd_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int PL = VEC * NV;
  constexpr int GROUPS = 32 / LANES;
  constexpr bool WIDE = (LANES == 32);

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
  const int LP = p.L * p.P;

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    // grad_out run owned by this lane (identical in every lane group)
    float g[PL];
    {
      const float* go = WIDE ? p.grad_out + (static_cast<size_t>(w.bq) * p.Hh + w.h) * p.C
                             : p.grad_out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim;
  #pragma unroll
      for (int j = 0; j < NV; ++j)
  #pragma unroll
        for (int i = 0; i < VEC; i += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(go + (j * LANES + sub) * VEC + i));
          g[j * VEC + i] = t.x; g[j * VEC + i + 1] = t.y; g[j * VEC + i + 2] = t.z; g[j * VEC + i + 3] = t.w;
        }
    }
    const float gws = (WIDE && p.grad_wsum != nullptr)
                          ? __ldg(p.grad_wsum + static_cast<size_t>(w.bq) * p.Hh + w.h) : 0.f;

    if (MODE == GD4D_MODE_C) {
      head_softmax(p, w, sw);
      gsum[lane] = 0.f;
      gsum[lane + 32] = 0.f;
      for (int i = lane; i < 3 * kMaxLP; i += 32) doff[i] = 0.f;
    }
    const int nvalid = build_candidates<MODE, CandB>(p, w, cands, false);

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
        alog = p.attn_logits + (static_cast<size_t>(w.bq) * p.N + n) * p.P * p.L + l;
        wt = 0.f;
        for (int pp = 0; pp < p.P; ++pp) wt += sigmoidf_(__ldg(alog + pp * p.L));
      }
      if (!active) wt = 0.f;
      const int W = p.level_w[l], H = p.level_h[l];
      const float ix = to_pixel(to_grid<MODE>(cu), static_cast<float>(W));
      const float iy = to_pixel(to_grid<MODE>(cv), static_cast<float>(H));
      const Footprint f = footprint(ix, iy, W, H);
      const VT* base = static_cast<const VT*>(p.value[l]);
      const size_t img = static_cast<size_t>(w.b) * p.N + n;
      const size_t e00 = ((img * H + f.y0) * W + f.x0) * p.C +
                         (WIDE ? 0 : static_cast<size_t>(w.h) * kHeadDim) + sub * VEC;
      const size_t rowst = static_cast<size_t>(W) * p.C;
      float c00[PL], c01[PL], c10[PL], c11[PL];
      const bool a00 = active & f.in00, a01 = active & f.in01, a10 = active & f.in10, a11 = active & f.in11;
      {
        uint4 r00[NV], r01[NV], r10[NV], r11[NV];
  #pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int o = j * LANES * VEC;
          r00[j] = ldg_nc_v4(base + e00 + o, a00);
          r01[j] = ldg_nc_v4(base + e00 + p.C + o, a01);
          r10[j] = ldg_nc_v4(base + e00 + rowst + o, a10);
          r11[j] = ldg_nc_v4(base + e00 + rowst + p.C + o, a11);
        }
  #pragma unroll
        for (int j = 0; j < NV; ++j) pin(r00[j], r01[j], r10[j], r11[j]);
  #pragma unroll
        for (int j = 0; j < NV; ++j) {
          Slice<VT>::unpack(r00[j], &c00[j * VEC]);
          Slice<VT>::unpack(r01[j], &c01[j * VEC]);
          Slice<VT>::unpack(r10[j], &c10[j * VEC]);
          Slice<VT>::unpack(r11[j], &c11[j * VEC]);
        }
      }
      const float w00 = (1.f - f.tx) * (1.f - f.ty), w01 = f.tx * (1.f - f.ty);
      const float w10 = (1.f - f.tx) * f.ty, w11 = f.tx * f.ty;

      // feature-map gradient: dL/df_c += wt * w_c * g   (vector reductions, no return value)
      float* gv = p.grad_value[l];
      if (gv != nullptr && wt != 0.f) {
        const float s00 = wt * w00, s01 = wt * w01, s10 = wt * w10, s11 = wt * w11;
  #pragma unroll
        for (int j = 0; j < NV; ++j)
  #pragma unroll
          for (int i = 0; i < VEC; i += 4) {
            const int o = j * LANES * VEC + i;
            const float g0 = g[j * VEC + i], g1 = g[j * VEC + i + 1], g2 = g[j * VEC + i + 2],
                        g3 = g[j * VEC + i + 3];
            if (a00) red_add_v4(gv + e00 + o, s00 * g0, s00 * g1, s00 * g2, s00 * g3);
            if (a01) red_add_v4(gv + e00 + p.C + o, s01 * g0, s01 * g1, s01 * g2, s01 * g3);
            if (a10) red_add_v4(gv + e00 + rowst + o, s10 * g0, s10 * g1, s10 * g2, s10 * g3);
            if (a11) red_add_v4(gv + e00 + rowst + p.C + o, s11 * g0, s11 * g1, s11 * g2, s11 * g3);
          }
      }

      float sdot = 0.f, dxdot = 0.f, dydot = 0.f;
  #pragma unroll
      for (int i = 0; i < PL; ++i) {
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
      if (WIDE) {  // the bias rides as an all-ones channel: 1 inside the map, 0 outside
        const float i00 = f.in00 ? 1.f : 0.f, i01 = f.in01 ? 1.f : 0.f, i10 = f.in10 ? 1.f : 0.f,
                    i11 = f.in11 ? 1.f : 0.f;
        sdot += gws * (w00 * i00 + w01 * i01 + w10 * i10 + w11 * i11);
        dxdot += gws * ((i01 - i00) * (1.f - f.ty) + (i11 - i10) * f.ty);
        dydot += gws * ((i10 - i00) * (1.f - f.tx) + (i11 - i01) * f.tx);
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
      float* ga = p.grad_attn_logits + (static_cast<size_t>(w.bq) * p.Hh + w.h) * LP;
      if (lane < LP) atomicAdd(ga + lane, s0 * (g0 - dot));
      if (lane + 32 < LP) atomicAdd(ga + lane + 32, s1 * (g1 - dot));
    }

    float rX = 0.f, rY = 0.f, rZ = 0.f;
    for (int k = lane; k < nvalid; k += 32) {
      const CandB cd = cands[k];
      const int n = cd.np >> 8;
      const int pi = cd.np & 0xff;
      const float* M = p.lidar2img + (static_cast<size_t>(w.b) * p.N + n) * 16;
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
          atomicAdd(p.grad_cam_logits + static_cast<size_t>(w.b) * p.N * p.Q +
                        static_cast<size_t>(n) * p.Q + w.q,
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
        float* gr = p.grad_ref + static_cast<size_t>(w.bq) * 3;
        atomicAdd(gr + 0, rX * p.pc_span[0]);
        atomicAdd(gr + 1, rY * p.pc_span[1]);
        atomicAdd(gr + 2, rZ * p.pc_span[2]);
      }
    }
    if (MODE == GD4D_MODE_C && p.grad_offsets != nullptr) {
      __syncwarp();
      float* go = p.grad_offsets + (static_cast<size_t>(w.bq) * p.Hh + w.h) * p.P * 3;
      for (int i = lane; i < p.P * 3; i += 32) atomicAdd(go + i, doff[i]);
    }
    __syncwarp();  // per-warp shared-memory state is reused by the next work item
  }
  work_end(p, wi);
}

template <int MODE, typename VT, int LANES, int NV>
static int launch_bwd(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  auto kern = xview_bwd_kernel<MODE, VT, LANES, NV>;
  if (g.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem);
    if (e != cudaSuccess) return GD4D_ERR_CUDA;
  }
  int grid = g.grid;
  if (p.sched != nullptr) {  // persistent grid: one resident wave, warps claim work dynamically
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, g.block, g.smem) != cudaSuccess)
      return GD4D_ERR_CUDA;
    const long long resident = static_cast<long long>(sms) * (occ > 0 ? occ : 1);
    if (resident < grid) grid = static_cast<int>(resident);
  }
  kern<<<grid, g.block, g.smem, stream>>>(p, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

int dispatch_backward(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream) {
  const bool bf16 = p.value_dtype == GD4D_BF16;
  if (p.mode == GD4D_MODE_A) {
    return bf16 ? launch_bwd<GD4D_MODE_A, __nv_bfloat16, 4, 1>(p, g, stream)
                : launch_bwd<GD4D_MODE_A, float, 8, 1>(p, g, stream);
  }
  if (p.mode == GD4D_MODE_C && !p.wide) {
    return bf16 ? launch_bwd<GD4D_MODE_C, __nv_bfloat16, 4, 1>(p, g, stream)
                : launch_bwd<GD4D_MODE_C, float, 8, 1>(p, g, stream);
  }
  if (p.mode == GD4D_MODE_C && p.wide) {
    if (bf16) return g.nv == 1 ? launch_bwd<GD4D_MODE_C, __nv_bfloat16, 32, 1>(p, g, stream)
                               : launch_bwd<GD4D_MODE_C, __nv_bfloat16, 32, 2>(p, g, stream);
    return g.nv == 1 ? launch_bwd<GD4D_MODE_C, float, 32, 1>(p, g, stream)
                     : launch_bwd<GD4D_MODE_C, float, 32, 2>(p, g, stream);
  }
  return GD4D_ERR_UNSUPPORTED;
}

}  // namespace gd4d
