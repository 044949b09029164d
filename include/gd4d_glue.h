/*
 * gd4d_glue.h -- C ABI of the small fused kernels that sit either side of the
 * cross-view sampling kernels inside one decoder layer.
 *
 * The reference spells these as chains of 5-20 one-line torch ops on (B*Q, 3) or
 * (B*Q, 256) tensors; at B*Q = 900 every one of them is a launch-latency-bound
 * kernel.  Each entry point below is ONE launch.  Same conventions as
 * gd4d_xview.h: plain C, raw fp32 DEVICE pointers, CUDA stream as void*, no
 * allocation, no synchronisation, 0 or a negative gd4d_status.
 *
 * Reference lines replaced (projects/mmdet3d_plugin/models/utils/):
 *   gd4d_inverse_sigmoid_fwd/bwd  inverse_sigmoid, detr3d_transformer.py:28-43 and
 *                                 deform3d_cross_attn.py:16-31 (the latter also clamps to max=1)
 *   gd4d_ref_update               reference-point refinement in logit space,
 *                                 detr3d_transformer.py:201-214
 *   gd4d_bias_act                 Linear bias (+ReLU) epilogue: FFN Linear-ReLU (mmcv FFN), the
 *                                 regression branches' Linear-ReLU (detr3d_head.py:72-95)
 *   gd4d_add_layernorm_fwd/bwd    "x (+ residual (+ pos_feat)) -> LayerNorm (-> ReLU)":
 *                                 the post-norm residual sums of the decoder layer
 *                                 (operation_order self_attn,norm,cross_attn,norm,ffn,norm;
 *                                 detr3d_transformer.py:386-390, deform3d_cross_attn.py:326-339)
 *                                 and the Linear-LN-ReLU stages of position_encoder
 *                                 (detr3d_transformer.py:297-304, deform3d_cross_attn.py:112-121)
 */
#ifndef GD4D_GLUE_H_
#define GD4D_GLUE_H_

#include "gd4d_xview.h"

#ifdef __cplusplus
extern "C" {
#endif

/* y = log(x1 / x2),  xc = clamp(x,0,1), x1 = clamp(xc, eps[, 1]), x2 = clamp(1-xc, eps[, 1]). */
GD4D_API int gd4d_inverse_sigmoid_fwd(const float* x, float* y, int64_t n, float eps,
                                      int32_t clamp_max, void* cuda_stream);
/* gx = gy * [0<=x<=1] * ([xc>=eps]/x1 + [1-xc>=eps]/x2)   (torch clamp passes gradient on ties) */
GD4D_API int gd4d_inverse_sigmoid_bwd(const float* x, const float* gy, float* gx, int64_t n,
                                      float eps, int32_t clamp_max, void* cuda_stream);

/* new_ref[r,0:2] = sigmoid(reg[r,0:2] + inverse_sigmoid(ref[r,0:2]))
 * new_ref[r,2]   = sigmoid(reg[r,4]   + inverse_sigmoid(ref[r,2]))
 * reg is (rows, reg_stride) with reg_stride >= 5 (the 10-wide box code); ref/new_ref (rows,3). */
GD4D_API int gd4d_ref_update(const float* reg, int32_t reg_stride, const float* ref, float* new_ref,
                             int64_t rows, float eps, void* cuda_stream);

/* y = [relu](y + bias) in place: the bias/activation epilogue of a Linear whose GEMM ran
 * without one (fp32 cuBLAS runs it as a separate kernel, and ReLU as another).
 * y (rows,C), bias (C), C % 4 == 0, 16-byte aligned. */
GD4D_API int gd4d_bias_act(float* y, const float* bias, int64_t rows, int32_t C, int32_t relu,
                           void* cuda_stream);

/* s = x + xbias + r1 + r2 (xbias (C) broadcast over rows; xbias, r1, r2 optional);
 * y = LayerNorm_C(s) * gamma + beta;  y = max(y,0) if relu;  y2 = y + pos (pos, y2 optional,
 * both or neither: the "query + query_pos" the next attention block starts with).
 * Writes y (rows,C), mean/rstd (rows) and s (s_out is required whenever s != x).
 * C must be a multiple of 128 and <= 1024; all row pointers 16-byte aligned. */
/* y_copy1 / y_copy2 (optional): identical copies of y, one per further autograd consumer of the result (the FFN
 * input and the next residual; the next layer's value input, its residual and the stacked output): each consumer
 * hands back its own gradient and gd4d_add_layernorm_bwd sums them in registers -- autograd would otherwise launch
 * one elementwise add per extra consumer (3 per decoder layer). */
GD4D_API int gd4d_add_layernorm_fwd(const float* x, const float* xbias, const float* r1,
                                    const float* r2, const float* gamma, const float* beta,
                                    const float* pos, float* y, float* y2, float* y_copy1, float* y_copy2,
                                    float* s_out, float* mean, float* rstd, int64_t rows, int32_t C, float eps,
                                    int32_t relu, void* cuda_stream);
/* gs = dL/ds for the incoming gradients gy, gy2 (of y2), g_copy1, g_copy2 (of the copies) -- any may be NULL, at
 * least one is not; the same gs
 * flows to x, r1 and r2.  With relu the incoming gradient is first masked by [y > 0] (recomputed
 * from s, mean, rstd, gamma, beta).  If g_masked != NULL the effective incoming gradient
 * (summed and/or masked) is written out for the deferred gamma/beta reduction. */
GD4D_API int gd4d_add_layernorm_bwd(const float* gy, const float* gy2, const float* g_copy1,
                                    const float* g_copy2, const float* s,
                                    const float* mean, const float* rstd, const float* gamma,
                                    const float* beta, float* gs, float* g_masked, int64_t rows,
                                    int32_t C, int32_t relu, void* cuda_stream);

/* AdamW (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected, step counter on
 * the device so the launch is CUDA-graph capturable) over MANY tensors in ONE launch.
 *   table_dev      device array of per-tensor records (all fp32, n elements each)
 *   block_map_dev  device array of n_blocks (tensor index, chunk index) int32 pairs: CTA b updates
 *                  elements [chunk*gd4d_adamw_chunk(), +gd4d_adamw_chunk()) of tensor table[idx]
 *   step_dev       device float: the step count AFTER increment (1 on the first update)
 * Replaces the optimizer step of the reference's training loop (mmcv OptimizerHook ->
 * torch.optim.AdamW, projects/configs/detr3d/detr3d_res50.py optimizer = dict(type='AdamW', ...)). */
typedef struct gd4d_adamw_tensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
} gd4d_adamw_tensor;

GD4D_API int gd4d_adamw_chunk(void);
GD4D_API int gd4d_adamw_multi(const gd4d_adamw_tensor* table_dev, const int32_t* block_map_dev,
                              int32_t n_blocks, const float* step_dev, float lr, float beta1,
                              float beta2, float eps, float weight_decay, void* cuda_stream);

/* Softmax backward over the last dim in one pass: grad_in = probs * (grad_out - sum_j grad_out_j probs_j).
 * rows x cols fp32, cols % 4 == 0, cols <= 1024, 16-byte aligned; grad_in may alias grad_out.  The softmax of
 * the decoder layer's self-attention (mmcv MultiheadAttention -> nn.MultiheadAttention, 8 x 900 x 900 per layer). */
GD4D_API int gd4d_softmax_bwd(const float* grad_out, const float* probs, float* grad_in, int64_t rows, int32_t cols,
                              void* cuda_stream);

/* fp32-accurate GEMM on the tensor cores (csrc/gemm_tf32x3.cu): C[M,N] = A . B^T (+ bias[N]) (relu), error-
 * compensated 3xTF32 (tcgen05.mma.kind::tf32, fp32 accumulate in tensor memory; relative error ~1e-6 like an
 * fp32 FMA chain).  Replaces the cuBLAS SIMT sgemm calls behind the reference's nn.Linear / F.linear sites of
 * the decoder layer (mmcv FFN / MultiheadAttention projections, deform3d_cross_attn.py:211-212,326,
 * detr3d_head.py:72-95) and their autograd.
 *   a_mn_major = 0: A is (M,K) row-major with row stride lda      (x in x.W^T; dY in dY.W)
 *   a_mn_major = 1: A is (K,M) row-major with row stride lda      (dY in dY^T.X)
 *   b_mn_major = 0: B is (N,K) row-major with row stride ldb      (W in x.W^T)
 *   b_mn_major = 1: B is (K,N) row-major with row stride ldb      (W in dY.W; X in dY^T.X)
 * C is (M,N) row-major with row stride ldc.  batch > 1: operand b of the batch starts stride_* floats further.
 * A, B 16-byte aligned; lda, ldb, stride_a, stride_b and each operand's contiguous extent multiples of 4
 * (else GD4D_ERR_ALIGN / GD4D_ERR_UNSUPPORTED: the caller keeps its library GEMM for such shapes). */
GD4D_API int gd4d_gemm_tf32x3(const float* A, int64_t lda, int32_t a_mn_major, const float* B, int64_t ldb,
                              int32_t b_mn_major, float* C, int64_t ldc, const float* bias, int32_t relu, int32_t M,
                              int32_t N, int32_t K, int32_t batch, int64_t stride_a, int64_t stride_b,
                              int64_t stride_c, void* cuda_stream);

/* Exact-fp32 (FFMA) GEMM for the small-M shapes of the decoder layer (csrc/sgemm_small.cu): same contract and
 * argument meaning as gd4d_gemm_tf32x3.  The default for M = B*Q = 900: a tcgen05.mma.kind::tf32 costs 153 cycles
 * whatever its N (tools/umma_latency.cu), so 8 row tiles of 128 cannot beat a register-tiled SIMT kernel there. */
GD4D_API int gd4d_sgemm_small(const float* A, int64_t lda, int32_t a_mn_major, const float* B, int64_t ldb,
                              int32_t b_mn_major, float* C, int64_t ldc, const float* bias, int32_t relu, int32_t M,
                              int32_t N, int32_t K, int32_t batch, int64_t stride_a, int64_t stride_b,
                              int64_t stride_c, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GD4D_GLUE_H_ */
