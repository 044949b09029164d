// extern "C" boundary: validation, launch geometry, dispatch.  No allocation, no
// synchronisation, stream-ordered, re-entrant (include/gd4d_xview.h).
#include "xview_common.cuh"

namespace gd4d {
int dispatch_forward(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream);
int dispatch_backward(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream);
int dispatch_backward_sorted(const gd4d_xview_params& p, const LaunchGeom& g, int stages, cudaStream_t stream);
long long sorted_ws_bytes(const gd4d_xview_params& p);
int dispatch_forward_tma(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream);
int dispatch_v2(const gd4d_xview_params& p, const LaunchGeom& g, cudaStream_t stream, bool backward);
int dispatch_pack(const void* src, void* dst, int src_dtype, int dst_dtype, int64_t images, int C,
                  int H, int W, cudaStream_t stream);

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int validate(const gd4d_xview_params* p, bool backward, LaunchGeom* g) {
  if (p == nullptr) return GD4D_ERR_NULL;
  if (p->abi_version != GD4D_ABI_VERSION) return GD4D_ERR_UNSUPPORTED;
  if (p->mode != GD4D_MODE_A && p->mode != GD4D_MODE_C && p->mode != GD4D_MODE_V2)
    return GD4D_ERR_UNSUPPORTED;
  if (p->value_dtype != GD4D_F32 && p->value_dtype != GD4D_BF16) return GD4D_ERR_UNSUPPORTED;
  if (p->B <= 0 || p->Q <= 0 || p->N <= 0 || p->Hh <= 0 || p->L <= 0 || p->P <= 0 || p->C <= 0)
    return GD4D_ERR_DIMS;
  if (p->L > GD4D_MAX_LEVELS || p->P > 255 || p->N >= (1 << 22)) return GD4D_ERR_DIMS;
  g->nv = 1;
  if (p->wide) {
    if (p->mode != GD4D_MODE_C) return GD4D_ERR_UNSUPPORTED;
    const long long row_bytes = static_cast<long long>(p->C) * (p->value_dtype == GD4D_BF16 ? 2 : 4);
    if (row_bytes != 512 && row_bytes != 1024) return GD4D_ERR_HEAD_DIM;
    g->nv = static_cast<int>(row_bytes / 512);
  } else {
    if (p->C % p->Hh != 0) return GD4D_ERR_DIMS;
    if (p->C / p->Hh != kHeadDim) return GD4D_ERR_HEAD_DIM;
  }
  if (!(p->img_h > 0.f) || !(p->img_w > 0.f)) return GD4D_ERR_DIMS;
  for (int l = 0; l < p->L; ++l) {
    if (p->level_h[l] <= 0 || p->level_w[l] <= 0) return GD4D_ERR_DIMS;
    if (p->value[l] == nullptr) return GD4D_ERR_NULL;
    if (!aligned16(p->value[l])) return GD4D_ERR_ALIGN;
    if (backward && p->grad_value[l] != nullptr && !aligned16(p->grad_value[l])) return GD4D_ERR_ALIGN;
  }
  if (p->ref == nullptr || p->lidar2img == nullptr || p->attn_logits == nullptr) return GD4D_ERR_NULL;
  if (p->mode == GD4D_MODE_C) {
    if (p->offsets == nullptr || p->cam_logits == nullptr) return GD4D_ERR_NULL;
    if (p->L * p->P > kMaxLP) return GD4D_ERR_UNSUPPORTED;
    if (p->gen_stride < 0 ||
        (p->gen_stride > 0 && p->gen_stride < p->Hh * p->L * p->P + p->Hh * p->P * 3 + p->N))
      return GD4D_ERR_DIMS;
  } else if (p->gen_stride != 0) {
    return GD4D_ERR_UNSUPPORTED;
  }
  if (p->mode == GD4D_MODE_V2) {
    if (p->offsets == nullptr) return GD4D_ERR_NULL;
    // the reference multiplies (..,L,P) weights with (..,P,L) samples: only L == P runs
    if (p->L != p->P || p->L * p->P > kMaxLP) return GD4D_ERR_UNSUPPORTED;
  }
  if (backward) {
    if (p->grad_out == nullptr) return GD4D_ERR_NULL;
    if (!aligned16(p->grad_out)) return GD4D_ERR_ALIGN;
  } else {
    if (p->out == nullptr) return GD4D_ERR_NULL;
    if (!aligned16(p->out)) return GD4D_ERR_ALIGN;
  }
  const long long warps = static_cast<long long>(p->B) * p->Q * p->Hh;
  const long long ctas = (warps + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 0x7fffffffLL) return GD4D_ERR_DIMS;
  const int pp = p->mode == GD4D_MODE_C ? p->P : 1;
  const long long cand_cap = static_cast<long long>(p->N) * pp;
  // backward keeps 16 more bytes per candidate (coordinate-gradient accumulators)
  const bool big = backward || p->mode == GD4D_MODE_V2;   // V2 uses the backward layout both ways
  const long long per_cand = big ? 32 : 16;
  // forward: 32 gather records of 48 B + softmax weights + candidates
  const long long recb = (backward && p->mode != GD4D_MODE_V2) ? 32 * 96 : 0;  // backward records
  const long long per_warp = big ? recb + sizeof(float) * kMaxLP * 5 + per_cand * cand_cap
                                 : 32 * 48 + sizeof(float) * kMaxLP + per_cand * cand_cap;
  const long long smem = per_warp * kWarpsPerCta;
  if (smem > 200 * 1024) return GD4D_ERR_UNSUPPORTED;
  g->grid = static_cast<int>(ctas);
  g->block = kWarpsPerCta * 32;
  g->smem = static_cast<int>(smem);
  g->cand_cap = static_cast<int>(cand_cap);
  return GD4D_OK;
}
}  // namespace gd4d

extern "C" {

int gd4d_abi_version(void) { return GD4D_ABI_VERSION; }

int gd4d_params_size(void) { return static_cast<int>(sizeof(gd4d_xview_params)); }

const char* gd4d_strerror(int status) {
  switch (status) {
    case GD4D_OK: return "ok";
    case GD4D_ERR_NULL: return "required pointer is NULL";
    case GD4D_ERR_DIMS: return "bad or inconsistent dimension";
    case GD4D_ERR_HEAD_DIM: return "unsupported channel width (narrow: C/Hh must be 32; wide: C*elem must be 512 or 1024 bytes)";
    case GD4D_ERR_ALIGN: return "pointer not 16-byte aligned";
    case GD4D_ERR_UNSUPPORTED: return "unsupported mode / dtype / size";
    case GD4D_ERR_CUDA: return "CUDA launch failed";
    default: return "unknown gd4d status";
  }
}

int gd4d_xview_launch_info(const gd4d_xview_params* p, int32_t* grid, int32_t* block,
                           int32_t* smem_bytes) {
  gd4d::LaunchGeom g{};
  const int st = gd4d::validate(p, false, &g);
  if (st != GD4D_OK) return st;
  if (grid) *grid = g.grid;
  if (block) *block = g.block;
  if (smem_bytes) *smem_bytes = g.smem;
  return GD4D_OK;
}

int gd4d_xview_forward(const gd4d_xview_params* p, void* cuda_stream) {
  gd4d::LaunchGeom g{};
  const int st = gd4d::validate(p, false, &g);
  if (st != GD4D_OK) return st;
  if (p->mode == GD4D_MODE_V2)
    return gd4d::dispatch_v2(*p, g, static_cast<cudaStream_t>(cuda_stream), false);
  if (p->mode == GD4D_MODE_C && p->wide && p->sched != nullptr && (p->flags & GD4D_FLAG_TMA_FORWARD))
    return gd4d::dispatch_forward_tma(*p, g, static_cast<cudaStream_t>(cuda_stream));
  return gd4d::dispatch_forward(*p, g, static_cast<cudaStream_t>(cuda_stream));
}

int gd4d_xview_backward(const gd4d_xview_params* p, void* cuda_stream) {
  gd4d::LaunchGeom g{};
  const int st = gd4d::validate(p, true, &g);
  if (st != GD4D_OK) return st;
  if (p->mode == GD4D_MODE_V2)
    return gd4d::dispatch_v2(*p, g, static_cast<cudaStream_t>(cuda_stream), true);
  if (p->mode == GD4D_MODE_C && p->wide && p->bwd_ws != nullptr)
    return gd4d::dispatch_backward_sorted(*p, g, (p->flags & GD4D_FLAG_BWD_PRESORTED) ? 4 :
                                                 (p->flags & GD4D_FLAG_BWD_EMITTED) ? 6 : 7,
                                          static_cast<cudaStream_t>(cuda_stream));
  if (p->flags & (GD4D_FLAG_BWD_PRESORTED | GD4D_FLAG_BWD_EMITTED)) return GD4D_ERR_UNSUPPORTED;
  return gd4d::dispatch_backward(*p, g, static_cast<cudaStream_t>(cuda_stream));
}

int gd4d_xview_backward_sort(const gd4d_xview_params* p, void* cuda_stream) {
  gd4d::LaunchGeom g{};
  gd4d_xview_params q;
  if (p == nullptr) return GD4D_ERR_NULL;
  q = *p;
  static float dummy_aligned[4] __attribute__((aligned(16)));
  if (q.grad_out == nullptr) q.grad_out = dummy_aligned;       // not read by the sort; validate() wants it non-NULL
  const int st = gd4d::validate(&q, true, &g);
  if (st != GD4D_OK) return st;
  if (p->mode != GD4D_MODE_C || !p->wide || p->bwd_ws == nullptr) return GD4D_ERR_UNSUPPORTED;
  return gd4d::dispatch_backward_sorted(*p, g, 3, static_cast<cudaStream_t>(cuda_stream));
}

int64_t gd4d_xview_bwd_ws_bytes(const gd4d_xview_params* p) {
  gd4d::LaunchGeom g{};
  const int st = gd4d::validate(p, false, &g);
  if (st != GD4D_OK && st != GD4D_ERR_NULL) return st;      // `out` may be unset: only the dimensions matter
  if (p == nullptr) return GD4D_ERR_NULL;
  if (p->mode != GD4D_MODE_C || !p->wide) return GD4D_ERR_UNSUPPORTED;
  const long long n = gd4d::sorted_ws_bytes(*p);
  return n < 0 ? GD4D_ERR_DIMS : n;
}

int gd4d_pack_nchw(const void* src, void* dst, int32_t src_dtype, int32_t dst_dtype, int64_t images,
                   int32_t C, int32_t H, int32_t W, void* cuda_stream) {
  if (src == nullptr || dst == nullptr) return GD4D_ERR_NULL;
  if (images <= 0 || C <= 0 || H <= 0 || W <= 0) return GD4D_ERR_DIMS;
  return gd4d::dispatch_pack(src, dst, src_dtype, dst_dtype, images, C, H, W,
                             static_cast<cudaStream_t>(cuda_stream));
}

int gd4d_unpack_nhwc(const float* src, float* dst, int64_t images, int32_t C, int32_t H, int32_t W,
                     void* cuda_stream) {
  if (src == nullptr || dst == nullptr) return GD4D_ERR_NULL;
  if (images <= 0 || C <= 0 || H <= 0 || W <= 0) return GD4D_ERR_DIMS;
  // the pack kernels are a batched 2-D transpose (rows x cols -> cols x rows): the inverse is the
  // same transpose with the roles of the channel axis and the pixel axis swapped
  return gd4d::dispatch_pack(src, dst, GD4D_F32, GD4D_F32, images, H * W, C, 1,
                             static_cast<cudaStream_t>(cuda_stream));
}

int gd4d_unpack_nhwc_cast(const float* src, void* dst, int32_t dst_dtype, int64_t images, int32_t C, int32_t H,
                          int32_t W, void* cuda_stream) {
  if (src == nullptr || dst == nullptr) return GD4D_ERR_NULL;
  if (images <= 0 || C <= 0 || H <= 0 || W <= 0) return GD4D_ERR_DIMS;
  if (dst_dtype != GD4D_F32 && dst_dtype != GD4D_BF16) return GD4D_ERR_UNSUPPORTED;
  return gd4d::dispatch_pack(src, dst, GD4D_F32, dst_dtype, images, H * W, C, 1, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
