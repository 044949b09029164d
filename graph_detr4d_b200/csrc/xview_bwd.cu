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
#include "xview_bwd_records.cuh"

namespace gd4d {

template <int MODE, typename VT, int LANES, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
xview_bwd_kernel(const __grid_constant__ gd4d_xview_params p, const int cand_cap) {
  constexpr int VEC = Slice<VT>::VEC;
  constexpr int PL = VEC * NV;
  constexpr int GROUPS = 32 / LANES;
  constexpr bool WIDE = (LANES == 32);
  constexpr int INF = 2;  // items in flight per group

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = lane / LANES;
  const int sub = lane % LANES;
  const size_t warp_bytes = sizeof(RecB) * 32 + sizeof(float) * kMaxLP * 5 + sizeof(CandB) * cand_cap;
  RecB* recs = reinterpret_cast<RecB*>(smem_raw + warp * warp_bytes);
  float* sw = reinterpret_cast<float*>(recs + 32);  // softmax weights
  float* gsum = sw + kMaxLP;                        // sum_n wcam*(s.g) per (l,p)
  float* doff = gsum + kMaxLP;                      // dL/d offset (p,3)
  CandB* cands = reinterpret_cast<CandB*>(doff + 3 * kMaxLP);
  const int LP = p.L * p.P;
  const bool l2_prefetch = (p.flags & GD4D_FLAG_L2_PREFETCH) != 0;

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    // grad_out run owned by this lane (identical in every lane group)
    float g[PL];
    {
      const float* go = WIDE ? p.grad_out + ((static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q) * p.C
                             : p.grad_out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(go + (j * LANES + sub) * VEC + i));
          g[j * VEC + i] = t.x; g[j * VEC + i + 1] = t.y; g[j * VEC + i + 2] = t.z; g[j * VEC + i + 3] = t.w;
        }
    }
    // The feature-gradient reductions use their own channel->lane map: vector r of a lane
    // sits at channel (r*LANES + sub)*4, so that every warp-wide RED instruction covers
    // LANES*16 CONTIGUOUS bytes of the fp32 grad row.  For fp32 maps that equals the load
    // layout; for bf16 maps (8 channels per 16-byte load) it does not, and reusing the
    // load layout would make each RED touch 16 of every 32 bytes (r1: 2x slower atomics).
    constexpr int NR = PL / 4;
    float gr[NR][4];
    {
      const float* go = WIDE ? p.grad_out + ((static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q) * p.C
                             : p.grad_out + static_cast<size_t>(w.bq) * p.C + w.h * kHeadDim;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (VEC == 4) {
          gr[r][0] = g[r * 4]; gr[r][1] = g[r * 4 + 1]; gr[r][2] = g[r * 4 + 2]; gr[r][3] = g[r * 4 + 3];
        } else {
          const float4 t = __ldg(reinterpret_cast<const float4*>(go + (r * LANES + sub) * 4));
          gr[r][0] = t.x; gr[r][1] = t.y; gr[r][2] = t.z; gr[r][3] = t.w;
        }
      }
    }
    const float gws = (WIDE && p.grad_wsum != nullptr)
                          ? __ldg(p.grad_wsum + (static_cast<size_t>(w.b) * p.Hh + w.h) * p.Q + w.q) : 0.f;

    if (MODE == GD4D_MODE_C) {
      head_softmax(p, w, sw);
      gsum[lane] = 0.f;
      gsum[lane + 32] = 0.f;
      for (int i = lane; i < 3 * kMaxLP; i += 32) doff[i] = 0.f;
    }
    const int nvalid = build_candidates<MODE, CandB>(p, w, cands, false);

    // ---- phase 2: re-gather, dot with grad_out, scatter feature gradients -------------------
    const int total = nvalid * p.L;
    for (int c0 = 0; c0 < total; c0 += 32) {
      recs[lane] = build_record_bwd<MODE, VT, WIDE>(p, cands, sw, c0 + lane, total, w);
      __syncwarp();
      const int nchunk = min(32, total - c0);
      for (int j0 = 0; j0 < nchunk; j0 += INF * GROUPS) {
        uint4 raw[INF][4][NV];
#pragma unroll
        for (int u = 0; u < INF; ++u) {
          const RecB* r = recs + (j0 + u * GROUPS + grp);
          const VT* base = static_cast<const VT*>(p.value[r->meta & 0xff]) + sub * VEC;
          const long long o00 = r->o00, o01 = r->o01, o10 = r->o10, o11 = r->o11;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const int o = j * LANES * VEC;
            raw[u][0][j] = ldg_nc_v4_all(reinterpret_cast<const char*>(base + o00 + o));
            raw[u][1][j] = ldg_nc_v4_all(reinterpret_cast<const char*>(base + o01 + o));
            raw[u][2][j] = ldg_nc_v4_all(reinterpret_cast<const char*>(base + o10 + o));
            raw[u][3][j] = ldg_nc_v4_all(reinterpret_cast<const char*>(base + o11 + o));
          }
        }
        if (l2_prefetch) {  // next batch's corner rows -> L2 while this batch's gathers are in flight
#pragma unroll
          for (int u = 0; u < INF; ++u) {
            const int it = j0 + (INF + u) * GROUPS + grp;
            if (it < nchunk) {
              const RecB* r = recs + it;
              const VT* base = static_cast<const VT*>(p.value[r->meta & 0xff]) + sub * VEC;
#pragma unroll
              for (int j = 0; j < NV; ++j) {
                const int o = j * LANES * VEC;
                prefetch_l2(base + r->o00 + o); prefetch_l2(base + r->o01 + o);
                prefetch_l2(base + r->o10 + o); prefetch_l2(base + r->o11 + o);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < INF; ++u)
#pragma unroll
          for (int j = 0; j < NV; ++j) pin(raw[u][0][j], raw[u][1][j], raw[u][2][j], raw[u][3][j]);

#pragma unroll
        for (int u = 0; u < INF; ++u) {
          const RecB* r = recs + (j0 + u * GROUPS + grp);
          const int meta = r->meta;
          const bool active = meta < 0;
          const int l = meta & 0xff;
          const float wt = r->wt;
          const float w00 = r->w00, w01 = r->w01, w10 = r->w10, w11 = r->w11;

          // feature-map gradient: dL/df_c += wt * w_c * g   (vector reductions, no return value)
          float* gv = p.grad_value[l];
          if (gv != nullptr && wt != 0.f) {
            gv += sub * 4;
            const float s00 = wt * w00, s01 = wt * w01, s10 = wt * w10, s11 = wt * w11;
            const long long o00 = r->o00, o01 = r->o01, o10 = r->o10, o11 = r->o11;
#pragma unroll
            for (int q = 0; q < NR; ++q) {
              const int o = q * LANES * 4;
              const float g0 = gr[q][0], g1 = gr[q][1], g2 = gr[q][2], g3 = gr[q][3];
              if (s00 != 0.f) red_add_v4(gv + o00 + o, s00 * g0, s00 * g1, s00 * g2, s00 * g3);
              if (s01 != 0.f) red_add_v4(gv + o01 + o, s01 * g0, s01 * g1, s01 * g2, s01 * g3);
              if (s10 != 0.f) red_add_v4(gv + o10 + o, s10 * g0, s10 * g1, s10 * g2, s10 * g3);
              if (s11 != 0.f) red_add_v4(gv + o11 + o, s11 * g0, s11 * g1, s11 * g2, s11 * g3);
            }
          }

          // per-corner dot products with grad_out, then three scalar combinations
          float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            float c00[VEC], c01[VEC], c10[VEC], c11[VEC];
            Slice<VT>::unpack(raw[u][0][j], c00);
            Slice<VT>::unpack(raw[u][1][j], c01);
            Slice<VT>::unpack(raw[u][2][j], c10);
            Slice<VT>::unpack(raw[u][3][j], c11);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              const float gi = g[j * VEC + i];
              d00 = fmaf(gi, c00[i], d00);
              d01 = fmaf(gi, c01[i], d01);
              d10 = fmaf(gi, c10[i], d10);
              d11 = fmaf(gi, c11[i], d11);
            }
          }
          float sdot = w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11;
          float dxdot = r->ax00 * d00 + r->ax01 * d01 + r->ax10 * d10 + r->ax11 * d11;  // W * ds/dix . g
          float dydot = r->ay00 * d00 + r->ay01 * d01 + r->ay10 * d10 + r->ay11 * d11;  // H * ds/diy . g
#pragma unroll
          for (int o = 1; o < LANES; o <<= 1) {
            sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
            dxdot += __shfl_xor_sync(0xffffffffu, dxdot, o);
            dydot += __shfl_xor_sync(0xffffffffu, dydot, o);
          }
          if (WIDE) {  // the bias rides as an all-ones channel: 1 inside the map, 0 outside
            sdot += gws * (w00 + w01 + w10 + w11);
            dxdot += gws * (r->ax00 + r->ax01 + r->ax10 + r->ax11);
            dydot += gws * (r->ay00 + r->ay01 + r->ay10 + r->ay11);
          }
          if (active && sub == 0) {
            const int k = (meta >> 16) & 0x7fff;
            atomicAdd(&cands[k].du, wt * dxdot);
            atomicAdd(&cands[k].dv, wt * dydot);
            if (MODE == GD4D_MODE_C) {
              atomicAdd(&gsum[(meta >> 8) & 0xff], r->cw * sdot);
              atomicAdd(&cands[k].cg, r->smw * sdot);
            } else if (p.grad_attn_logits != nullptr) {
              const int n = cands[k].np >> 8;
              const size_t ao = (static_cast<size_t>(w.bq) * p.N + n) * p.P * p.L + l;
              for (int pp = 0; pp < p.P; ++pp) {
                const float sg = sigmoidf_(__ldg(p.attn_logits + ao + pp * p.L));
                atomicAdd(p.grad_attn_logits + ao + pp * p.L, sg * (1.f - sg) * sdot);
              }
            }
          }
        }
      }
      __syncwarp();  // records are rebuilt by the next chunk
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
      float* ga = p.grad_attn_logits + attn_row_off(p, w);
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
          atomicAdd(p.grad_cam_logits + cam_off(p, w, n), cd.w * (1.f - cd.w) * cd.cg);
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
      float* go = p.grad_offsets + offsets_row_off(p, w);
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
