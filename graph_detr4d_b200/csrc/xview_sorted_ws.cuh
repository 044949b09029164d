// Scratch layout of the sorted wide backward (xview_bwd_sorted.cu), shared with the forward kernel, which can
// emit the contribution records itself (GD4D_FLAG_FWD_EMIT).
#pragma once
#include "xview_common.cuh"

namespace gd4d {

struct SortedWs {
  unsigned* counters;   // [0] slots handed out by K1 (4 per item), [1] number of sorted contributions,
                        // [2] K2 blocks done
  int* row_count;       // R    histogram; zero on entry, re-zeroed by K4
  int* row_start;       // R    block-local exclusive prefix
  int* block_sums;      // nblk exclusive prefix of the 2048-row block totals
  int* base;            // B*Q*Hh  first cid of each (b,q,head)
  int4* rec;            // cap  32-byte records (two int4): {value row ptr, grad-map row ptr | go_row, coef, row | -1, rank}
  int4* sorted;         // cap  32-byte SRec records (two int4 each), ordered by pixel row
  float* dots;          // cap  value_row . grad_out_row per contribution
  long long level_row0[GD4D_MAX_LEVELS + 1];   // first global row of each level (+ total)
  int R, nblk;
  long long cap;
};

// carve p.bwd_ws; returns the bytes needed (negative: dimensions out of range).  ws / base_ptr may be NULL.
long long sorted_ws_layout(const gd4d_xview_params& p, SortedWs* ws, char* base_ptr);

// One corner contribution, as the emit stage writes it: two int4
//   lo = {value-row pointer (2 x 32 bits), element offset of the row in its level's fp32 grad map, level}
//   hi = {grad_out row, coefficient bits (wt * bilinear corner weight), global pixel row | -1 (out of the map), rank}
__device__ __forceinline__ void emit_contribution(const SortedWs& ws, int cid, bool in_map, const void* vrow, long long goff,
                                                  int level, int row, int go_row, float coef) {
  int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, -1, 0);
  if (in_map) {
    const unsigned long long vp = reinterpret_cast<unsigned long long>(vrow);
    lo = make_int4(static_cast<int>(vp), static_cast<int>(vp >> 32), static_cast<int>(goff), level);
    hi = make_int4(go_row, __float_as_int(coef), row, atomicAdd(ws.row_count + row, 1));
  }
  ws.rec[2 * cid] = lo;
  ws.rec[2 * cid + 1] = hi;
}

}  // namespace gd4d
