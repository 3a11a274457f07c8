/*
 * elo_b200.h -- C ABI of libelo_b200.so, the B200 (sm_100a) implementation of EfficientLO-Net's
 * projection-aware point-cloud hot path.  Plain pointers and sizes only; no torch / TensorFlow
 * types.  Every pointer is a DEVICE pointer unless the function name ends in `_host`.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, as the reference uses).
 * All functions return 0 on success, a positive cudaError_t if the CUDA runtime failed, or a
 * negative ELO_ERR_* code; elo_last_error() gives the message (thread-local).
 * Nothing here allocates device memory, synchronises, or reads results back on the host, so
 * every device-pointer entry point is CUDA-graph capturable.
 *
 * Reference interfaces replaced (file:line relative to the reference repository):
 *   elo_fused_conv_select_k  <-  FusedConvSelectKLauncher  tf_ops/2d_conv_select_k/fused_conv.cpp:73,
 *                                                           tf_ops/2d_conv_select_k/fused_conv_g.cu:215
 *   elo_fused_conv_random_k  <-  FusedConvRandomKLauncher  tf_ops/2d_conv_random_k/fused_conv.cpp:73,
 *                                                           tf_ops/2d_conv_random_k/fused_conv_g.cu:162
 *   (+ the zero-fill of fused_conv.cpp:154-166, which the kernels here do themselves: every output
 *    element is written, callers need not clear the buffers.)
 */
#ifndef ELO_B200_H
#define ELO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ELO_OK 0
#define ELO_ERR_INVALID_ARGUMENT (-1) /* the conditions fused_conv.cpp:79-100 rejects, null pointers */
#define ELO_ERR_UNSUPPORTED (-2)      /* legal for the reference op but outside this build's limits */

const char *elo_last_error(void);
int elo_version(void);

/*
 * Projection-aware neighbour search.  Same parameter list and meaning as the reference Launchers,
 * plus the stream.  Shapes (C-contiguous):
 *   xyz1 (B,H,W,3) f32   xyz2 (B,small_h,small_w,3) f32   idx_n2 (B,npoints,2) i32 [h,w]
 *   random_hw (kH*kW) i32
 *   selected_bhw_idx (B,npoints,K,3) i32 [b,h,w]          selected_mask (B,npoints,K,1) f32
 *   valid_idx, valid_in_dis_idx (B,npoints,kH*kW,1) f32   -- may be NULL (no caller in the model
 *   reads them, utils/pointnet_util.py:49,106,197,272); they are 83 % of the op's bytes.
 * Limits: kH*kW <= 5000 and K <= 5000 (the reference's local-array size, fused_conv_g.cu:42-43),
 * H, W, small_h, small_w < 32768.
 */
int elo_fused_conv_select_k(int batch_size, int H, int W, int npoints, int kernel_size_H,
                            int kernel_size_W, int K, int flag_copy, float distance, int stride_h,
                            int stride_w, const float *xyz1, const float *xyz2, const int *idx_n2,
                            const int *random_hw, int *selected_bhw_idx, float *valid_idx,
                            float *valid_in_dis_idx, float *selected_mask, int small_h, int small_w,
                            void *stream);

int elo_fused_conv_random_k(int batch_size, int H, int W, int npoints, int kernel_size_H,
                            int kernel_size_W, int K, int flag_copy, float distance, int stride_h,
                            int stride_w, const float *xyz1, const float *xyz2, const int *idx_n2,
                            const int *random_hw, int *selected_bhw_idx, float *valid_idx,
                            float *valid_in_dis_idx, float *selected_mask, int small_h, int small_w,
                            void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ELO_B200_H */
