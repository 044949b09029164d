/*
 * gd4d_fpe.h -- C ABI of the feature position-embedding block of Graph-DETR4D's PE head
 * (second half):
 *
 *   Detr3DHeadPE.forward              projects/mmdet3d_plugin/models/dense_heads/detr3d_head_pe.py:519-553
 *   SELayer.forward                   projects/mmdet3d_plugin/models/dense_heads/detr3d_head_pe.py:239-243
 *   SinePositionalEncoding3D.forward  projects/mmdet3d_plugin/models/utils/positional_encoding.py:58-100
 *
 * Per FPN level the reference (a) builds a full-resolution (B,N,pad_h,pad_w) padding mask on the host
 * side loop and nearest-interpolates it to the level, (b) runs three cumsums, ~12 elementwise ops, two
 * stacks, a cat and a permute to get the 3F-channel sine embedding, and (c) after the 1x1 convolutions
 * (library code) combines everything with four more full-size elementwise passes:
 *
 *   pe   = fpe(pe, feat)  = pe * sigmoid(conv_expand(relu(conv_reduce(feat))))        :545, :239-243
 *   feat = feat + (pe + adapt_pos3d(sine))                                            :549-553
 *
 * Here: one launch writes a level's mask, one launch writes its sine embedding straight from the
 * per-camera image sizes (the mask is the complement of a rectangle, so the cumsums are closed-form),
 * one launch does the combine, one its backward.  The convolutions stay library calls.
 * fp32, the reference's op order (explicit round-to-nearest ops where the argument of sin/cos is
 * formed, so masked rows/columns -- whose normalised coordinate is (0 - 0.5)/eps*scale ~ -3e6 --
 * get bit-identical arguments).  All pointers are DEVICE pointers; same conventions as gd4d_xview.h.
 */
#ifndef GD4D_FPE_H_
#define GD4D_FPE_H_

#include "gd4d_xview.h"

#ifdef __cplusplus
extern "C" {
#endif

/* img_hw (B*N, 2) int32 = img_metas[b]['img_shape'][n][0:2] (unpadded rows, cols).
 * mask (B*N, H, W) uint8: 1 where the level pixel's nearest source pixel
 *   (F.interpolate 'nearest': src = min(floor(dst * float(pad)/float(size)), pad-1)) lies outside the
 *   camera's image, i.e. the reference's interpolated padding mask (:523-536). */
GD4D_API int gd4d_level_mask(const int32_t* img_hw, uint8_t* mask, int32_t BN, int32_t H, int32_t W,
                             int32_t pad_h, int32_t pad_w, void* cuda_stream);

/* out (B*N, 3*F, H, W) fp32 = SinePositionalEncoding3D(mask) for the mask above, channels
 * [n-embedding F | y-embedding F | x-embedding F], sin on even / cos on odd channels.
 * dim_t (F) fp32 = temperature ** (2 * (i // 2) / F) as torch computes it (passed in so that it is
 * bit-identical to the reference's tensor); normalize != 0: e = (e + offset) / (e_last + eps) * scale. */
GD4D_API int gd4d_sine_pe3d(const int32_t* img_hw, const float* dim_t, float* out, int32_t B, int32_t N,
                            int32_t H, int32_t W, int32_t pad_h, int32_t pad_w, int32_t F,
                            int32_t normalize, float scale, float eps, float offset, void* cuda_stream);

/* out = feat + (pe * sigmoid(gate) + sine), n elements each (any layout, all four the same). */
GD4D_API int gd4d_fpe_combine_fwd(const float* feat, const float* pe, const float* gate, const float* sine,
                                  float* out, int64_t n, void* cuda_stream);
/* backward of the combine: grad_pe = g * sigmoid(gate); grad_gate = g * pe * s * (1 - s).
 * (grad_feat = grad_sine = g: the caller aliases them.)  Either output may be NULL. */
GD4D_API int gd4d_fpe_combine_bwd(const float* grad_out, const float* pe, const float* gate, float* grad_pe,
                                  float* grad_gate, int64_t n, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GD4D_FPE_H_ */
