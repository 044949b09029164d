/*
 * gd4d_xview.h -- C ABI of the B200-native cross-view 3D->2D feature-sampling
 * attention (Graph-DETR4D decoder hot path).
 *
 * Plain C, raw DEVICE pointers and sizes, a CUDA stream passed as void*; no
 * torch types, no allocation, no device synchronisation, re-entrant.  Every
 * entry point returns 0 on success or a negative gd4d_status; nothing throws
 * across the boundary.  The shared library (libgd4d_xview.so) is built from
 * graph_detr4d_b200/csrc/ for sm_100a only.
 *
 * Each entry point replaces one piece of the reference (paths relative to the
 * reference repo, projects/mmdet3d_plugin/models/utils/):
 *
 *   gd4d_xview_forward   mode A : feature_sampling + sigmoid*mask + 3 sums
 *                                 detr3d_transformer.py:376-383, 397-438
 *                        mode C : 3D graph-offset points -> projection -> mask
 *                                 -> softmax*mask -> MultiScaleDeformableAttn
 *                                 -> per-camera sigmoid weight -> sum over cams
 *                                 deform3d_cross_attn.py:227-258, 274, 281-284,
 *                                 301-304, 320-324   (mmcv ms_deform_attn_forward)
 *                        mode V2: per-(cam,head,level,point) 2D offsets
 *                                 detr3d_transformer.py:602-627, 636-709
 *   gd4d_xview_backward  what autograd + mmcv ms_deform_attn_backward +
 *                        grid_sampler_2d_backward do for the same lines
 *   gd4d_pack_nchw       the per-layer flatten/transpose/cat of the feature maps
 *                        deform3d_cross_attn.py:264-269 (done ONCE per forward here)
 */
#ifndef GD4D_XVIEW_H_
#define GD4D_XVIEW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GD4D_API __attribute__((visibility("default")))
#else
#define GD4D_API
#endif

#define GD4D_ABI_VERSION 6
#define GD4D_MAX_LEVELS 8

typedef enum gd4d_status {
  GD4D_OK = 0,
  GD4D_ERR_NULL = -1,        /* required pointer is NULL                       */
  GD4D_ERR_DIMS = -2,        /* non-positive / inconsistent dimension          */
  GD4D_ERR_HEAD_DIM = -3,    /* narrow: C/Hh != 32; wide: C*elem not 512/1024 B  */
  GD4D_ERR_ALIGN = -4,       /* a feature/out pointer is not 16-byte aligned   */
  GD4D_ERR_UNSUPPORTED = -5, /* unknown mode / dtype / L*P > 64 / N*P too big  */
  GD4D_ERR_CUDA = -6         /* launch failed (cudaGetLastError != success)    */
} gd4d_status;

typedef enum gd4d_mode {
  GD4D_MODE_A = 0,  /* Detr3DCrossAtten   (centre point, sigmoid weights per cam) */
  GD4D_MODE_C = 1,  /* Deform3DCrossAttn  (3D offset points, softmax, cam weight) */
  GD4D_MODE_V2 = 2  /* Detr3DCrossAttenV2 (2D offsets, softmax per cam & head)    */
} gd4d_mode;

typedef enum gd4d_dtype { GD4D_F32 = 0, GD4D_BF16 = 1 } gd4d_dtype;

/* forward, wide mode, sched set: stage the corner rows in shared memory with TMA bulk
 * copies (cp.async.bulk + mbarrier pipeline) instead of register gathers */
#define GD4D_FLAG_TMA_FORWARD 1u
/* forward and backward: prefetch the NEXT batch's corner rows into L2 (prefetch.global.L2) while
 * the current batch's register gathers are in flight */
#define GD4D_FLAG_L2_PREFETCH 2u
/* measurement only (bench.py times the owner kernel of the sorted backward as call-with minus call-without):
 * run the sorted backward WITHOUT its owner pass -- the feature gradient is not produced and the small gradients
 * are meaningless */
#define GD4D_FLAG_BWD_SKIP_OWNER 4u
/* backward, sorted path: bwd_ws already holds this call's sorted contribution records (gd4d_xview_backward_sort ran
 * on the same scratch with the same forward inputs): run only the owner and finish kernels */
#define GD4D_FLAG_BWD_PRESORTED 8u
/* forward, mode C wide, bwd_ws set: the forward kernel also EMITS the sorted backward's contribution records into
 * bwd_ws (it builds the same per-item records anyway); the matching backward call passes GD4D_FLAG_BWD_EMITTED and
 * the same scratch and skips its emit kernel.  One scratch per in-flight (forward, backward) pair. */
#define GD4D_FLAG_FWD_EMIT 16u
#define GD4D_FLAG_BWD_EMITTED 32u

/*
 * One decoder-layer invocation.  All pointers are device pointers.
 *
 * Feature maps ("value") are CHANNEL-LAST, one base pointer per FPN level:
 *   value[l] : (B*N, level_h[l], level_w[l], C)   fp32 or bf16, 16-byte aligned
 * with camera index n fastest inside the leading B*N axis (image = b*N + n),
 * i.e. exactly torch's channels_last storage of the reference's
 * feat.view(B*N, C, H, W), and exactly mmcv's (bs, num_keys, heads, dims) value
 * layout taken level by level.
 *
 * Per-mode tensors (fp32, contiguous):
 *            ref (B,Q,3) in [0,1]; lidar2img (B,N,4,4) row-major
 *   mode A : attn_logits (B,Q,N,P,L)              offsets = cam_logits = NULL
 *   mode C : attn_logits (B,Q,Hh,L,P)             offsets (B,Q,Hh,P,3) metres
 *            cam_logits  (B, Q*N) raw Linear output; the kernel indexes it as the
 *            reference's .view(B,N,Q,1) does: weight(n,q) = flat[n*Q + q]
 *   mode V2: attn_logits (B,Q,N,Hh,L,P)           offsets (B,Q,N,Hh,L,P,2) pixels
 * Activations (sigmoid / softmax) are applied inside the kernel.
 *
 * Outputs:  out (B,Q,C) fp32.
 *           mask (optional, may be NULL) uint8: A/V2 (B,Q,N); C (B,N,Q,Hh,P).
 *
 * "Wide" sampling (mode C, wide = 1) -- gather-then-project.  The reference runs
 * value_proj (a 256x256 Linear) over EVERY pixel of every camera in every layer
 * (deform3d_cross_attn.py:278) and then samples head h's 32-channel slice.  By
 * linearity  sum_s w_s * (W f_s + b) = W (sum_s w_s f_s) + b * sum_s w_s, so with
 * wide = 1 each head samples all C channels of the RAW feature maps:
 *   out  (B,Hh,Q,C) = sum_s w_s * bilinear(f_s)          (fp32, head-major)
 *   wsum (B,Hh,Q)   = sum_s w_s * (in-bounds corner weight sum)   [multiplies the bias]
 * and the caller applies W_v's head slice to `out` with one tiny batched GEMM.
 * No dense per-layer GEMM, no per-layer copy of the maps, and in backward the
 * feature gradient lands directly in ONE grad map shared by all decoder layers.
 * grad_out is then (B,Hh,Q,C) and grad_wsum (B,Hh,Q).  C must be 32 lanes * 16 B
 * * {1,2}: fp32 C in {128,256}, bf16 C in {256,512}.
 *
 * Backward (gd4d_xview_backward) reads grad_out (B,Q,C) and ACCUMULATES with
 * atomics into caller-zeroed buffers; any grad pointer may be NULL to skip it:
 *   grad_value[l] fp32 channel-last, same shape as value[l]
 *   grad_attn_logits / grad_offsets / grad_cam_logits / grad_ref : as the inputs
 */
typedef struct gd4d_xview_params {
  int32_t abi_version;          /* must be GD4D_ABI_VERSION */
  int32_t mode;                 /* gd4d_mode  */
  int32_t value_dtype;          /* gd4d_dtype */
  int32_t B, Q, N, Hh, L, P, C;
  int32_t wide;                 /* mode C only: 0 = head slices of projected value, 1 = see above */
  uint32_t flags;               /* GD4D_FLAG_* */
  int32_t gen_stride;           /* mode C only.  0: attn_logits / offsets / cam_logits (and their
                                   gradients) are the dense tensors described above.  > 0: they are
                                   COLUMN BLOCKS of one row-major (B*Q, gen_stride) fp32 matrix -- the
                                   output of ONE GEMM over the three generator Linears' concatenated
                                   weights: each pointer addresses its block's first column in row 0,
                                   row b*Q+q holds query q's Hh*L*P logits, Hh*P*3 offsets and N camera
                                   logits; the camera logit of (n,q) is element e = n*Q+q of the (Q,N)
                                   block, i.e. row e/N, column e%N (the reference's view quirk). */
  int32_t level_h[GD4D_MAX_LEVELS];
  int32_t level_w[GD4D_MAX_LEVELS];
  float pc_lo[3];               /* pc_range[0:3]                              */
  float pc_span[3];             /* float32(pc_range[3+i] - pc_range[i])       */
  float img_h, img_w;           /* img_metas[0]['img_shape'][0][0:2], unpadded */
  const void* value[GD4D_MAX_LEVELS];
  const float* value_bias;      /* optional (C): bias of a value_proj folded into
                                   the sampler: out += bias * sum(w * inbounds) */
  const float* ref;
  const float* lidar2img;
  const float* attn_logits;
  const float* offsets;
  const float* cam_logits;
  uint32_t* sched;              /* optional: 2 uint32, zeroed ONCE by the caller.  When set the
                                   kernels run a persistent grid whose warps claim (b,q,head)
                                   work items from this counter (no wave-quantisation tail) and
                                   the last warp to finish resets it, so it is reusable by the
                                   next launch ON THE SAME STREAM without another memset.       */
  float* out;
  float* wsum;                  /* wide only */
  uint8_t* mask;
  /* backward only */
  const float* grad_out;
  const float* grad_wsum;       /* wide only, may be NULL (treated as zeros) */
  float* grad_value[GD4D_MAX_LEVELS];
  float* grad_value_bias;
  float* grad_attn_logits;
  float* grad_offsets;
  float* grad_cam_logits;
  float* grad_ref;
  /* backward, mode C wide only: scratch of at least gd4d_xview_bwd_ws_bytes() bytes, 256-byte aligned,
   * ZERO on first use (the launch leaves its counters and histogram zeroed again, so the same scratch
   * serves every later backward ON THE SAME STREAM without a memset).  When set, the backward sorts the
   * corner contributions by pixel row and lets one warp own each run (csrc/xview_bwd_sorted.cu): one
   * reduction per distinct row per 32 contributions instead of one per corner read.  NULL: the
   * atomics-per-corner kernel (csrc/xview_bwd.cu).  Same results up to fp32 summation order. */
  void* bwd_ws;
  int64_t bwd_ws_bytes;
} gd4d_xview_params;

GD4D_API int gd4d_abi_version(void);
/* sizeof(gd4d_xview_params) as compiled, so FFI mirrors can check their layout */
GD4D_API int gd4d_params_size(void);
GD4D_API const char* gd4d_strerror(int status);

/* bytes of dynamic shared memory / CTAs the forward launch will use (for tests
 * and the roofline bookkeeping); negative status on invalid params. */
GD4D_API int gd4d_xview_launch_info(const gd4d_xview_params* p, int32_t* grid, int32_t* block,
                           int32_t* smem_bytes);

GD4D_API int gd4d_xview_forward(const gd4d_xview_params* p, void* cuda_stream);
GD4D_API int gd4d_xview_backward(const gd4d_xview_params* p, void* cuda_stream);
/* The sort stage of the sorted backward alone (emit, scan, scatter into p->bwd_ws): needs only the FORWARD inputs
 * (value maps, ref, logits, offsets, lidar2img), not grad_out nor the grad maps, so a caller can run it right after
 * the forward on another stream and later call gd4d_xview_backward with GD4D_FLAG_BWD_PRESORTED on the same scratch.
 * One scratch per in-flight (forward, backward) pair. */
GD4D_API int gd4d_xview_backward_sort(const gd4d_xview_params* p, void* cuda_stream);
/* bytes of bwd_ws the sorted backward needs for these dimensions (mode C, wide); negative status otherwise */
GD4D_API int64_t gd4d_xview_bwd_ws_bytes(const gd4d_xview_params* p);

/* NCHW (images, C, H, W) -> channel-last (images, H, W, C) with optional cast.
 * src_dtype/dst_dtype: gd4d_dtype (fp32->fp32, fp32->bf16, bf16->bf16). */
GD4D_API int gd4d_pack_nchw(const void* src, void* dst, int32_t src_dtype, int32_t dst_dtype,
                   int64_t images, int32_t C, int32_t H, int32_t W, void* cuda_stream);

/* Inverse of gd4d_pack_nchw for fp32: channel-last (images, H, W, C) -> NCHW (images, C, H, W).
 * Hands the shared feature-gradient map back to an NCHW producer (the FPN) in one tiled
 * transpose instead of a strided elementwise copy. */
GD4D_API int gd4d_unpack_nhwc(const float* src, float* dst, int64_t images, int32_t C, int32_t H,
                     int32_t W, void* cuda_stream);

/* gd4d_unpack_nhwc with the cast folded in: fp32 channel-last gradient map -> NCHW in dst_dtype (gd4d_dtype), for
 * producers that hold their maps in bf16 (autograd wants the gradient in the leaf's dtype and layout). */
GD4D_API int gd4d_unpack_nhwc_cast(const float* src, void* dst, int32_t dst_dtype, int64_t images, int32_t C,
                                   int32_t H, int32_t W, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GD4D_XVIEW_H_ */
