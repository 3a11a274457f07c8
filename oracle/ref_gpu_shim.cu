/*
 * oracle/ref_gpu_shim.cu -- TEST INFRASTRUCTURE.  extern "C" doors onto the reference's own
 * Launchers (tf_ops/2d_conv_select_k/fused_conv_g.cu:215, tf_ops/2d_conv_random_k/fused_conv_g.cu:162),
 * whose .cu files oracle/Makefile compiles UNMODIFIED for sm_100a from /root/reference.
 * Reproduces the op wrapper's four cudaMemset calls (fused_conv.cpp:154-166).  Device pointers.
 */
#include <cuda_runtime.h>

void FusedConvSelectKLauncher(int, int, int, int, int, int, int, int, float, int, int, const float *,
                              const float *, const int *, const int *, int *, float *, float *,
                              float *, int, int);
void FusedConvRandomKLauncher(int, int, int, int, int, int, int, int, float, int, int, const float *,
                              const float *, const int *, const int *, int *, float *, float *,
                              float *, int, int);

template <typename L>
static int run(L launcher, int B, int H, int W, int n, int kH, int kW, int K, int flag_copy,
               float distance, int sh, int sw, const float *xyz1, const float *xyz2,
               const int *idx_n2, const int *random_hw, int *idx, float *valid, float *vdis,
               float *mask, int h2, int w2, int zero_fill)
{
    if (zero_fill) {
        cudaMemset(idx, 0, sizeof(int) * (size_t)B * n * K * 3);
        cudaMemset(valid, 0, sizeof(float) * (size_t)B * n * kH * kW);
        cudaMemset(vdis, 0, sizeof(float) * (size_t)B * n * kH * kW);
        cudaMemset(mask, 0, sizeof(float) * (size_t)B * n * K);
    }
    launcher(B, H, W, n, kH, kW, K, flag_copy, distance, sh, sw, xyz1, xyz2, idx_n2, random_hw, idx,
             valid, vdis, mask, h2, w2);
    return (int)cudaGetLastError();
}

extern "C" int ref_gpu_select_k(int B, int H, int W, int n, int kH, int kW, int K, int flag_copy,
                                float distance, int sh, int sw, const float *xyz1,
                                const float *xyz2, const int *idx_n2, const int *random_hw, int *idx,
                                float *valid, float *vdis, float *mask, int h2, int w2, int zero_fill)
{
    return run(FusedConvSelectKLauncher, B, H, W, n, kH, kW, K, flag_copy, distance, sh, sw, xyz1,
               xyz2, idx_n2, random_hw, idx, valid, vdis, mask, h2, w2, zero_fill);
}

extern "C" int ref_gpu_random_k(int B, int H, int W, int n, int kH, int kW, int K, int flag_copy,
                                float distance, int sh, int sw, const float *xyz1,
                                const float *xyz2, const int *idx_n2, const int *random_hw, int *idx,
                                float *valid, float *vdis, float *mask, int h2, int w2, int zero_fill)
{
    return run(FusedConvRandomKLauncher, B, H, W, n, kH, kW, K, flag_copy, distance, sh, sw, xyz1,
               xyz2, idx_n2, random_hw, idx, valid, vdis, mask, h2, w2, zero_fill);
}
