// Sorted ("owner computes") backward of the wide cross-view sampling kernel (mode C, wide = 1).
//
// Why: the atomics backward (xview_bwd.cu) issues one 1 KB vector reduction PER CORNER READ into the
// shared fp32 grad map -- 486 k reductions = 497 MB through the L2 atomic units per launch at N = 6,
// 54 of them on the same row at the coarsest level.  ncu (profiles/r2_xview_kernels_ncu_full.csv) and
// the L2 roofs (profiles/l2_peaks.json: red.add.v4.f32 4.6-5.6 TB/s) put that kernel at 0.83 of the
// L2-atomic roof and 0.36 of HBM.  But only 96 k DISTINCT rows are touched (5x reuse; 2x at level 0,
// 54x at level 3).  So: sort the corner contributions by pixel row and let one warp own a run of equal
// rows -- it loads the value row ONCE, gathers grad_out rows (7 MB, L2-resident) for the run's
// contributions, forms the dot products the small gradients need AND the row's feature gradient in
// registers, and retires the row with ONE reduction.  Reductions drop 486 k -> ~110 k, value-row
// gathers 486 k -> ~110 k.
//
// Five stream-ordered launches (all through gd4d_xview_backward when p.bwd_ws is set; the first three -- the sort --
// need only the forward's inputs and can run early, on another stream, through gd4d_xview_backward_sort):
//   K1 emit      warp per (b,q,head): candidates + records exactly as xview_bwd.cu builds them; every
//                in-map corner takes a slot in its row's histogram (atomicAdd returns its rank) and
//                writes a 32-byte record {value-row pointer, grad-map offset + level | grad_out row,
//                wt * w_corner, row, rank} at cid = base + item*4 + corner.  With GD4D_FLAG_FWD_EMIT the
//                FORWARD kernel writes the same records (xview_fwd.cu) and this launch is skipped.
//   K2 scan      exclusive prefix over the row histogram (block-local + last-block-done carry scan)
//   K3 scatter   record -> its sorted position row_start[row] + rank (pure permutation); re-zeroes the
//                histogram for the next call
//   K4 owner     warp per 32 consecutive sorted contributions (perfectly balanced; a row that spans
//                warps is simply reduced by each): one record per lane, value row fetched at run starts
//                only (cp.async ring, one batch ahead), 4 grad_out rows in flight, dot -> dots[cid],
//                acc += coef * g, red.add.v4.f32 at run ends
//   K5 finish    warp per (b,q,head), ONE LANE PER ITEM: rebuilds the same records, reads its 4 dots
//                and does what the tail of xview_bwd.cu does (softmax / sigmoid / projection chain)
// Results equal xview_bwd.cu up to fp32 summation order (tests/test_xview_gpu.py runs both).
#include "xview_common.cuh"
#include "xview_bwd_records.cuh"
#include "xview_sorted_ws.cuh"

namespace gd4d {

constexpr int kScanBlock = 2048;   // rows per K2 block (256 threads x 8)

long long sorted_ws_layout(const gd4d_xview_params& p, SortedWs* ws, char* base_ptr) {
  long long R = 0;
  for (int l = 0; l < p.L; ++l) {
    if (ws) ws->level_row0[l] = R;
    R += static_cast<long long>(p.B) * p.N * p.level_h[l] * p.level_w[l];
  }
  if (R > 0x7ffffff0LL) return -1;
  for (int l = 0; l < p.L; ++l)   // grad-map element offsets travel as int32
    if (static_cast<long long>(p.B) * p.N * p.level_h[l] * p.level_w[l] * p.C > 0x7ffffff0LL) return -1;
  const long long items = static_cast<long long>(p.B) * p.Q * p.Hh * p.N * p.P * p.L;
  const long long cap = items * 4;
  if (cap > 0x7ffffff0LL) return -1;
  const long long nblk = (R + kScanBlock - 1) / kScanBlock;
  long long off = 0;
  auto take = [&](long long bytes) { const long long o = off; off += (bytes + 255) / 256 * 256; return o; };
  const long long o_cnt = take(16), o_rc = take(R * 4), o_rs = take(R * 4), o_bs = take(nblk * 4);
  const long long o_base = take(static_cast<long long>(p.B) * p.Q * p.Hh * 4);
  const long long o_rec = take(cap * 32), o_sorted = take(cap * 32), o_dots = take(cap * 4);
  if (ws) {
    ws->level_row0[p.L] = R;
    ws->R = static_cast<int>(R); ws->nblk = static_cast<int>(nblk); ws->cap = cap;
    ws->counters = reinterpret_cast<unsigned*>(base_ptr + o_cnt);
    ws->row_count = reinterpret_cast<int*>(base_ptr + o_rc);
    ws->row_start = reinterpret_cast<int*>(base_ptr + o_rs);
    ws->block_sums = reinterpret_cast<int*>(base_ptr + o_bs);
    ws->base = reinterpret_cast<int*>(base_ptr + o_base);
    ws->rec = reinterpret_cast<int4*>(base_ptr + o_rec);
    ws->sorted = reinterpret_cast<int4*>(base_ptr + o_sorted);
    ws->dots = reinterpret_cast<float*>(base_ptr + o_dots);
  }
  return off;
}

long long sorted_ws_bytes(const gd4d_xview_params& p) { return sorted_ws_layout(p, nullptr, nullptr); }

// ------------------------------------------------------------------------------------------------
// K1 emit / K5 finish: one kernel body, warp per (b, q, head), one LANE per (candidate, level) item
// ------------------------------------------------------------------------------------------------
template <typename VT, bool FINISH>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
xview_bwd_items_kernel(const __grid_constant__ gd4d_xview_params p, const __grid_constant__ SortedWs ws,
                       const int cand_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const size_t warp_bytes = sizeof(float) * kMaxLP * 5 + sizeof(CandB) * cand_cap;
  float* sw = reinterpret_cast<float*>(smem_raw + warp * warp_bytes);   // softmax weights
  float* gsum = sw + kMaxLP;                                              // sum_n wcam*(s.g) per (l,p)
  float* doff = gsum + kMaxLP;                                            // dL/d offset (p,3)
  CandB* cands = reinterpret_cast<CandB*>(doff + 3 * kMaxLP);
  const int LP = p.L * p.P;
  if (FINISH && blockIdx.x == 0 && threadIdx.x == 0) {   // K3 / K4 are done with them: ready for the next call
    ws.counters[0] = 0u; ws.counters[1] = 0u; ws.counters[2] = 0u;
  }

  WorkIter wi;
  work_begin(p, wi);
  WarpCtx w;
  while (work_next(p, wi, w)) {
    head_softmax(p, w, sw);
    if (FINISH) {
      gsum[lane] = 0.f;
      gsum[lane + 32] = 0.f;
      for (int i = lane; i < 3 * kMaxLP; i += 32) doff[i] = 0.f;
    }
    const int nvalid = build_candidates<GD4D_MODE_C, CandB>(p, w, cands, false);
    const int total = nvalid * p.L;
    const int gw = w.bq * p.Hh + w.h;
    int base;
    if (FINISH) {
      base = ws.base[gw];
    } else {
      unsigned v = 0;
      if (lane == 0 && total > 0) v = atomicAdd(ws.counters, static_cast<unsigned>(4 * total));
      base = static_cast<int>(__shfl_sync(0xffffffffu, v, 0));
      if (lane == 0) ws.base[gw] = base;
    }
    const int go_row = (w.b * p.Hh + w.h) * p.Q + w.q;
    const float gws = (FINISH && p.grad_wsum != nullptr) ? __ldg(p.grad_wsum + go_row) : 0.f;

    for (int item = lane; item < total; item += 32) {
      const RecB r = build_record_bwd<GD4D_MODE_C, VT, true>(p, cands, sw, item, total, w);
      const int l = r.meta & 0xff;
      const long long z = (reinterpret_cast<const char*>(g_zero_row) - static_cast<const char*>(p.value[l])) /
                          static_cast<long long>(sizeof(VT));
      const long long o[4] = {r.o00, r.o01, r.o10, r.o11};
      const int cid0 = base + item * 4;
      if (!FINISH) {
        const float cw[4] = {r.wt * r.w00, r.wt * r.w01, r.wt * r.w10, r.wt * r.w11};
        const VT* vbase = static_cast<const VT*>(p.value[l]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool in_map = o[j] != z;
          emit_contribution(ws, cid0 + j, in_map, vbase + o[j], o[j], l,
                            in_map ? static_cast<int>(ws.level_row0[l] + o[j] / p.C) : -1, go_row, cw[j]);
        }
      } else {
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = (o[j] != z) ? ws.dots[cid0 + j] : 0.f;
        float sdot = r.w00 * d[0] + r.w01 * d[1] + r.w10 * d[2] + r.w11 * d[3];
        float dxdot = r.ax00 * d[0] + r.ax01 * d[1] + r.ax10 * d[2] + r.ax11 * d[3];   // W * ds/dix . g
        float dydot = r.ay00 * d[0] + r.ay01 * d[1] + r.ay10 * d[2] + r.ay11 * d[3];   // H * ds/diy . g
        // the bias rides as an all-ones channel: 1 inside the map, 0 outside
        sdot += gws * (r.w00 + r.w01 + r.w10 + r.w11);
        dxdot += gws * (r.ax00 + r.ax01 + r.ax10 + r.ax11);
        dydot += gws * (r.ay00 + r.ay01 + r.ay10 + r.ay11);
        const int k = (r.meta >> 16) & 0x7fff;
        atomicAdd(&cands[k].du, r.wt * dxdot);
        atomicAdd(&cands[k].dv, r.wt * dydot);
        atomicAdd(&gsum[(r.meta >> 8) & 0xff], r.cw * sdot);
        atomicAdd(&cands[k].cg, r.smw * sdot);
      }
    }
    __syncwarp();
    if (FINISH) {
      if (p.grad_attn_logits != nullptr) {   // softmax backward: dlogit_j = sm_j * (G_j - sum_k sm_k G_k)
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
        atomicAdd(&doff[pi * 3 + 0], dX);
        atomicAdd(&doff[pi * 3 + 1], dY);
        atomicAdd(&doff[pi * 3 + 2], dZ);
        if (p.grad_cam_logits != nullptr)
          atomicAdd(p.grad_cam_logits + cam_off(p, w, n), cd.w * (1.f - cd.w) * cd.cg);
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
      if (p.grad_offsets != nullptr) {
        __syncwarp();
        float* go = p.grad_offsets + offsets_row_off(p, w);
        for (int i = lane; i < p.P * 3; i += 32) atomicAdd(go + i, doff[i]);
      }
    }
    __syncwarp();  // per-warp shared-memory state is reused by the next work item
  }
  work_end(p, wi);
}

// ------------------------------------------------------------------------------------------------
// K2: exclusive prefix of the row histogram
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xview_bwd_scan_kernel(const SortedWs ws) {
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * kScanBlock + tid * 8;
  int v[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = (r0 + i < ws.R) ? ws.row_count[r0 + i] : 0;
    sum += v[i];
  }
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int wbase = 0;
  for (int i = 0; i < warp; ++i) wbase += warp_tot[i];
  int run = wbase + inc - sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (r0 + i < ws.R) ws.row_start[r0 + i] = run;
    run += v[i];
  }
  if (tid == 255) {
    ws.block_sums[blockIdx.x] = run;             // block total
    __threadfence();
    last = atomicAdd(ws.counters + 2, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // the last block to finish turns the block totals into exclusive prefixes (256 at a time)
  __threadfence();
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < ws.nblk; b0 += 256) {
    const int i = b0 + tid;
    const int x = i < ws.nblk ? *(volatile int*)(ws.block_sums + i) : 0;
    int s = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    int wb = carry_s;
    for (int k = 0; k < warp; ++k) wb += warp_tot[k];
    if (i < ws.nblk) ws.block_sums[i] = wb + s - x;
    __syncthreads();
    if (tid == 255) carry_s = wb + s;
    __syncthreads();
  }
  if (tid == 0) ws.counters[1] = static_cast<unsigned>(carry_s);   // number of sorted contributions
}

// ------------------------------------------------------------------------------------------------
// K3: scatter every contribution record (K1 already resolved its value-row pointer, grad-map row pointer --
//     NULL when no feature gradient is wanted -- grad_out row and coefficient) to its sorted position, adding
//     its cid.  Also returns the histogram entries to zero for the next call.
// ------------------------------------------------------------------------------------------------
struct __align__(16) SRec {
  const void* vrow;   // value row of this contribution's pixel
  int goff;           // element offset of the same row in the level's fp32 grad map (resolved by the owner: the
  int level;          //   grad map need not exist yet when the sort runs)
  int go_row;         // grad_out row (b, head, q)
  float coef;         // wt * bilinear corner weight
  int cid;            // where its dot product goes
  int row;            // global pixel row (run key)
};
static_assert(sizeof(SRec) == 32, "two 16-byte halves per record");

__global__ void __launch_bounds__(256) xview_bwd_scatter_kernel(const SortedWs ws) {
  const unsigned n = ws.counters[0];
  for (unsigned cid = blockIdx.x * blockDim.x + threadIdx.x; cid < n; cid += gridDim.x * blockDim.x) {
    const int4 hi = ws.rec[2 * cid + 1];                               // {go_row, coef, row | -1, rank}
    if (hi.z < 0) continue;
    const int4 lo = ws.rec[2 * cid];
    const int pos = ws.row_start[hi.z] + ws.block_sums[hi.z / kScanBlock] + hi.w;
    ws.sorted[2 * pos] = lo;
    ws.sorted[2 * pos + 1] = make_int4(hi.x, hi.y, static_cast<int>(cid), hi.z);   // SRec: go_row, coef, cid, row
    if (hi.w == 0) ws.row_count[hi.z] = 0;                             // one writer per row (K2 was the last reader)
  }
}

// ------------------------------------------------------------------------------------------------
// K4: owner pass.  A warp owns 32 consecutive sorted contributions (one record per lane, fields broadcast by
// shuffle), processed 4 at a time:
//   * value rows are fetched at run starts only, by cp.async into a per-warp shared-memory ring, one batch
//     AHEAD (no registers, no stall at a run start even where every contribution is its own run: level 0)
//   * the batch's 4 grad_out rows (L2-resident, 7 MB) are register gathers issued back to back
//   * per contribution: dot(value row, g) and acc += coef * g; the 4 dots are reduced together (butterfly
//     with halving payload: 6 shuffles per batch instead of 20) and lane u stores dot u
//   * at a run end the row's gradient leaves with one red.global.add.v4.f32 per lane vector
// ------------------------------------------------------------------------------------------------
constexpr int kOwnerWarps = 8;
constexpr int kOwnerU = 4;             // contributions per batch
constexpr int kOwnerSlots = 2 * kOwnerU;   // value-row ring: the current batch's and the next batch's run starts

template <typename VT, int NV>
__global__ void __launch_bounds__(kOwnerWarps * 32, 2)
xview_bwd_owner_kernel(const __grid_constant__ gd4d_xview_params p, const __grid_constant__ SortedWs ws) {
  constexpr int VEC = Slice<VT>::VEC;     // channels per 16-byte value load
  constexpr int PL = VEC * NV;            // channels per lane
  constexpr int GV = VEC / 4;             // float4 per value vector in the fp32 grad_out / grad-map rows
  constexpr int U = kOwnerU;
  constexpr int kRowBytes = 512 * NV;     // one value row (C * sizeof(VT))
  extern __shared__ __align__(16) unsigned char ring_all[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* ring = ring_all + wid * (kOwnerSlots * kRowBytes);
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring));
  const int n = static_cast<int>(ws.counters[1]);
  const int nchunks = (n + 31) / 32;
  const int warps = gridDim.x * kOwnerWarps;
  const SRec* sorted = reinterpret_cast<const SRec*>(ws.sorted);

  for (int c = blockIdx.x * kOwnerWarps + wid; c < nchunks; c += warps) {
    const int i0 = c * 32;
    const int cnt = min(32, n - i0);
    // one record per lane
    SRec me;
    me.vrow = nullptr; me.goff = 0; me.level = 0; me.go_row = 0; me.coef = 0.f; me.cid = 0; me.row = -1;
    if (lane < cnt) {
      const int4 lo = __ldg(reinterpret_cast<const int4*>(sorted + i0 + lane));
      const int4 hi = __ldg(reinterpret_cast<const int4*>(sorted + i0 + lane) + 1);
      me.vrow = reinterpret_cast<const void*>((static_cast<unsigned long long>(static_cast<unsigned>(lo.y)) << 32) |
                                              static_cast<unsigned>(lo.x));
      me.goff = lo.z; me.level = lo.w;
      me.go_row = hi.x; me.coef = __int_as_float(hi.y); me.cid = hi.z; me.row = hi.w;
    }
    const int prev = __shfl_up_sync(0xffffffffu, me.row, 1);
    const unsigned starts = __ballot_sync(0xffffffffu, lane < cnt && (lane == 0 || me.row != prev));

    // value rows of the run starts inside [u0, u0+U) -> ring slots (u0 / U & 1) * U + (i - u0); one commit per batch
    auto prefetch = [&](int u0) {
      if (u0 < cnt) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = u0 + u;
          if ((starts >> i) & 1u) {                                     // warp-uniform
            const unsigned long long vp = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(me.vrow), i);
            const uint32_t dst = ring_s + (((u0 / U) & 1) * U + u) * kRowBytes;
#pragma unroll
            for (int j = 0; j < NV; ++j)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (j * 32 + lane) * 16),
                           "l"(vp + (j * 32 + lane) * 16) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[PL], v[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) { acc[i] = 0.f; v[i] = 0.f; }
    unsigned long long cur_gv = 0;           // grad-map row of the open run (0: none / not wanted)
    bool open = false;

    auto flush = [&]() {
      if (open && cur_gv != 0) {
        float* gv = reinterpret_cast<float*>(cur_gv);
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int k = 0; k < GV; ++k)
            red_add_v4(gv + (j * 32 + lane) * VEC + k * 4, acc[j * VEC + k * 4], acc[j * VEC + k * 4 + 1],
                       acc[j * VEC + k * 4 + 2], acc[j * VEC + k * 4 + 3]);
      }
    };

    prefetch(0);
    for (int u0 = 0; u0 < cnt; u0 += U) {
      prefetch(u0 + U);                                                  // next batch's value rows
      // this batch's grad_out rows
      float4 g[U][NV][GV];
      float coef[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = min(u0 + u, 31);
        const int go = __shfl_sync(0xffffffffu, me.go_row, i);
        coef[u] = __shfl_sync(0xffffffffu, me.coef, i);                  // 0 for lanes past cnt
        const float* grow = p.grad_out + static_cast<long long>(go) * p.C;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int k = 0; k < GV; ++k) {
            const uint4 t = ldg_nc_v4_all(reinterpret_cast<const char*>(grow + (j * 32 + lane) * VEC + k * 4));
            g[u][j][k] = make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z),
                                     __uint_as_float(t.w));
          }
      }
      asm volatile("cp.async.wait_group 1;" ::: "memory");              // this batch's value rows have landed
      __syncwarp();
      float dot[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = u0 + u;
        if ((starts >> i) & 1u) {                                       // warp-uniform: a new run
          flush();
          {
            float* gl = p.grad_value[__shfl_sync(0xffffffffu, me.level, i)];
            const int goff = __shfl_sync(0xffffffffu, me.goff, i);
            cur_gv = gl != nullptr ? reinterpret_cast<unsigned long long>(gl + goff) : 0ull;
          }
          open = true;
          const unsigned char* slot = ring + (((u0 / U) & 1) * U + u) * kRowBytes;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const uint4 raw = *reinterpret_cast<const uint4*>(slot + (j * 32 + lane) * 16);
            Slice<VT>::unpack(raw, v + j * VEC);
          }
#pragma unroll
          for (int q = 0; q < PL; ++q) acc[q] = 0.f;
        }
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int k = 0; k < GV; ++k) {
            const float4 gg = g[u][j][k];
            const int b = j * VEC + k * 4;
            d0 = fmaf(v[b], gg.x, d0);     d1 = fmaf(v[b + 1], gg.y, d1);
            d0 = fmaf(v[b + 2], gg.z, d0); d1 = fmaf(v[b + 3], gg.w, d1);
            acc[b] = fmaf(coef[u], gg.x, acc[b]);         acc[b + 1] = fmaf(coef[u], gg.y, acc[b + 1]);
            acc[b + 2] = fmaf(coef[u], gg.z, acc[b + 2]); acc[b + 3] = fmaf(coef[u], gg.w, acc[b + 3]);
          }
        dot[u] = d0 + d1;
      }
      __syncwarp();                                                      // the ring slots of this batch are free again
      // 4 dots x 32 lanes -> lane u holds dot u: halve the payload at each of the first two butterfly steps
      {
        const bool up16 = (lane & 16) != 0;
        float s0 = up16 ? dot[0] : dot[2], s1 = up16 ? dot[1] : dot[3];          // send the pair the partner keeps
        float k0 = up16 ? dot[2] : dot[0], k1 = up16 ? dot[3] : dot[1];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);                              // lanes <16: dots 0,1; >=16: dots 2,3
        const bool up8 = (lane & 8) != 0;
        const float s = up8 ? k0 : k1;
        float k = up8 ? k1 : k0;                                                 // bit 3 picks the odd dot of the pair
        k += __shfl_xor_sync(0xffffffffu, s, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        // lane's dot index: 2 * bit4 + bit3; lanes 0, 8, 16, 24 write dots 0..3
        const int u = ((lane >> 4) << 1) | ((lane >> 3) & 1);
        const int cid = __shfl_sync(0xffffffffu, me.cid, min(u0 + u, 31));
        if ((lane & 7) == 0 && u0 + u < cnt) ws.dots[cid] = k;
      }
    }
    flush();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }
}

template <typename VT>
static int launch_sorted(const gd4d_xview_params& p, const LaunchGeom& g, const SortedWs& ws, int stages,
                         cudaStream_t stream) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return GD4D_ERR_CUDA;
  const int block = kWarpsPerCta * 32;
  const int smem = static_cast<int>((sizeof(float) * kMaxLP * 5 + sizeof(CandB) * g.cand_cap) * kWarpsPerCta);
  auto emit = xview_bwd_items_kernel<VT, false>;
  auto finish = xview_bwd_items_kernel<VT, true>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(emit, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(finish, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return GD4D_ERR_CUDA;
  }
  // emit / finish are chains of dependent round trips (project -> softmax -> records -> atomics), ~17 items on
  // 32 lanes: a persistent grid claiming work items adds two more round trips per item and leaves most of the
  // machine's warp slots empty.  One warp per (b, q, head), statically, all of them resident at once.
  gd4d_xview_params ps = p;
  ps.sched = nullptr;
  if (stages & 1) emit<<<g.grid, block, smem, stream>>>(ps, ws, g.cand_cap);   // needs the forward's inputs only
  if (stages & 2) {
    xview_bwd_scan_kernel<<<ws.nblk, 256, 0, stream>>>(ws);
    xview_bwd_scatter_kernel<<<sms * 16, 256, 0, stream>>>(ws);
  }
  if (!(stages & 4)) return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
  if (!(p.flags & GD4D_FLAG_BWD_SKIP_OWNER)) {
    const int osmem = kOwnerWarps * kOwnerSlots * 512 * g.nv;          // per-warp value-row rings
    auto own1 = xview_bwd_owner_kernel<VT, 1>;
    auto own2 = xview_bwd_owner_kernel<VT, 2>;
    if (osmem > 48 * 1024) {
      if (cudaFuncSetAttribute(g.nv == 1 ? own1 : own2, cudaFuncAttributeMaxDynamicSharedMemorySize, osmem) != cudaSuccess)
        return GD4D_ERR_CUDA;
    }
    if (g.nv == 1) own1<<<sms * 2, kOwnerWarps * 32, osmem, stream>>>(p, ws);
    else own2<<<sms * 2, kOwnerWarps * 32, osmem, stream>>>(p, ws);
  }
  finish<<<g.grid, block, smem, stream>>>(ps, ws, g.cand_cap);
  return cudaGetLastError() == cudaSuccess ? GD4D_OK : GD4D_ERR_CUDA;
}

// stages (bits): 1 = emit, 2 = scan + scatter, 4 = owner + finish.  7 = the whole backward; 3 = the sort alone
// (gd4d_xview_backward_sort); 4 = on a scratch sorted earlier (GD4D_FLAG_BWD_PRESORTED); 6 = on records the
// FORWARD kernel emitted (GD4D_FLAG_FWD_EMIT on the forward call, GD4D_FLAG_BWD_EMITTED here)
int dispatch_backward_sorted(const gd4d_xview_params& p, const LaunchGeom& g, int stages, cudaStream_t stream) {
  if (p.mode != GD4D_MODE_C || !p.wide || p.bwd_ws == nullptr) return GD4D_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(p.bwd_ws) & 255u) != 0) return GD4D_ERR_ALIGN;
  SortedWs ws;
  const long long need = sorted_ws_layout(p, &ws, static_cast<char*>(p.bwd_ws));
  if (need < 0) return GD4D_ERR_DIMS;
  if (p.bwd_ws_bytes < need) return GD4D_ERR_DIMS;
  return p.value_dtype == GD4D_BF16 ? launch_sorted<__nv_bfloat16>(p, g, ws, stages, stream)
                                    : launch_sorted<float>(p, g, ws, stages, stream);
}

}  // namespace gd4d
