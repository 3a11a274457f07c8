/*
 * oracle/ref_host_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile the reference's CUDA kernel
 * *bodies* (tf_ops/2d_conv_{select,random}_k/fused_conv_g.cu, everything above the Launcher) as
 * plain host C++, straight from /root/reference (oracle/Makefile pipes the source through sed
 * into g++; nothing of the reference is copied into this repository).
 *
 * Pre-included with  g++ -include ref_host_shim.h -DREF_KERNEL=<kernel> -DREF_RUNNER=<export>.
 * The runner plays the role of the op wrapper + launcher: it zero-fills the outputs
 * (fused_conv.cpp:154-166) and then "launches" <<<batch_size, T>>> by looping blockIdx.x / threadIdx.x.
 *
 * Caveat (also in DESIGN.md): host std::max(NaN, x) differs from the device FMNMX, and the host
 * build only contracts a*a+b*b+c*c to FMAs when g++ is given -mfma -ffp-contract=fast, so the GPU
 * build (libref_gpu.so) stays the canonical bit-exact reference; this one pins the restatement
 * on GPU-less machines and serves as the "reference" CPU baseline.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
using std::max;
#define __global__
struct RefDim3 { int x, y, z; };
static thread_local RefDim3 blockIdx, threadIdx, blockDim;

void REF_KERNEL(int batch_size, int H, int W, int npoints, int kernel_size_H, int kernel_size_W,
                int K, int flag_copy, float distance, int stride_h, int stride_w,
                const float *xyz1, const float *xyz2, const int *idx_n2, const int *random_hw,
                int *selected_bhw_idx, float *valid_idx, float *valid_in_dis_idx,
                float *selected_mask, int small_h, int small_w);

extern "C" int REF_RUNNER(int batch_size, int H, int W, int npoints, int kernel_size_H,
                          int kernel_size_W, int K, int flag_copy, float distance, int stride_h,
                          int stride_w, const float *xyz1, const float *xyz2, const int *idx_n2,
                          const int *random_hw, int *selected_bhw_idx, float *valid_idx,
                          float *valid_in_dis_idx, float *selected_mask, int small_h, int small_w,
                          int block_threads, int omp_threads)
{
    const size_t kt = (size_t)kernel_size_H * kernel_size_W;
    std::memset(selected_bhw_idx, 0, sizeof(int) * (size_t)batch_size * npoints * K * 3);
    std::memset(valid_idx, 0, sizeof(float) * (size_t)batch_size * npoints * kt);
    std::memset(valid_in_dis_idx, 0, sizeof(float) * (size_t)batch_size * npoints * kt);
    std::memset(selected_mask, 0, sizeof(float) * (size_t)batch_size * npoints * K);
    if (block_threads < 1) block_threads = 1;
    if (omp_threads < 1) omp_threads = 1;
    const long total = (long)batch_size * block_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(omp_threads)
    for (long i = 0; i < total; ++i) {
        blockIdx.x = (int)(i / block_threads);
        threadIdx.x = (int)(i % block_threads);
        blockDim.x = block_threads;
        REF_KERNEL(batch_size, H, W, npoints, kernel_size_H, kernel_size_W, K, flag_copy, distance,
                   stride_h, stride_w, xyz1, xyz2, idx_n2, random_hw, selected_bhw_idx, valid_idx,
                   valid_in_dis_idx, selected_mask, small_h, small_w);
    }
    return 0;
}
