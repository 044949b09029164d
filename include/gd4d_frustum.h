/*
 * gd4d_frustum.h -- C ABI of the fused frustum position-embedding input kernel
 * (SURVEY.md section 8f, row f4): the elementwise part of
 *
 *   Detr3DHeadPE.position_embeding    projects/mmdet3d_plugin/models/dense_heads/detr3d_head_pe.py:427-491
 *
 * up to, not including, the position_encoder 1x1 convolutions (library code).  The reference
 * builds a (B,N,W,H,D,4) frustum tensor, repeats the 4x4 img2lidar matrices over it, runs a
 * batched matmul and ~25 more full-size elementwise ops per FPN level, every forward.  Here one
 * launch per level writes the conv input directly:
 *
 *   out[bn, d*3+c, h, w] = inverse_sigmoid((img2lidar[bn] . [x*max(z,eps), y*max(z,eps), z, 1])[c] - lo[c]) / span[c])
 *       x = w*pad_w/W, y = h*pad_h/H, z = depth_start + bin_size*d*(d+1)          (:439-455)
 *   mask_out[b,n,h,w] = mask_in | (#{(d,c): coord outside [0,1]} > D/2)            (:476-478)
 *
 * fp32, the reference's op order, no FMA contraction in the projection; all DEVICE pointers
 * except pc_lo_span (6 host floats, read at launch).  Same conventions as gd4d_xview.h.
 */
#ifndef GD4D_FRUSTUM_H_
#define GD4D_FRUSTUM_H_

#include "gd4d_xview.h"

#ifdef __cplusplus
extern "C" {
#endif

/* img2lidar (B*N,16) fp32 = float32(inv(float64 lidar2img)); mask_in (B*N,H,W) uint8 or NULL;
 * out (B*N, 3*D, H, W) fp32; mask_out (B*N,H,W) uint8 (may be NULL);
 * pc_lo_span = {pc_range[0..2], float32(pc_range[3+i]-pc_range[i])};
 * bin_size = float32((pc_range[3]-depth_start) / (D*(1+D))). */
GD4D_API int gd4d_frustum_pe(const float* img2lidar, const uint8_t* mask_in, float* out,
                             uint8_t* mask_out, int32_t BN, int32_t H, int32_t W, int32_t D,
                             float pad_h, float pad_w, float depth_start, float bin_size,
                             const float* pc_lo_span, void* cuda_stream);

/* All FPN levels of one forward in ONE launch (the per-level call above is this with num_levels = 1).
 * mask_in / out / mask_out: HOST arrays of num_levels DEVICE pointers (mask_in, mask_out or single entries may
 * be NULL); level_h / level_w: HOST arrays.  num_levels <= 8.  position_embeding's per-level loop :427-491. */
GD4D_API int gd4d_frustum_pe_levels(const float* img2lidar, const uint8_t* const* mask_in, float* const* out,
                                    uint8_t* const* mask_out, int32_t BN, int32_t num_levels,
                                    const int32_t* level_h, const int32_t* level_w, int32_t D, float pad_h,
                                    float pad_w, float depth_start, float bin_size, const float* pc_lo_span,
                                    void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GD4D_FRUSTUM_H_ */
