// tc_test.cu -- single dense layer on the tcgen05 path (elo_tc.cuh), used by tests/test_tc_gpu.py to pin
// the TMEM / descriptor conventions against a float64 matmul before the fused kernels rely on them.
//   Y[128 x N] = act(X[128 x K] * W[K x N] + bias),  K % 16 == 0, K <= 192, N in {64, 128}
#include <cuda_runtime.h>

#include "../../include/elo_b200.h"
#include "../../efficientlo-net_b200/csrc/elo_common.cuh"
#include "../../efficientlo-net_b200/csrc/elo_mlp.cuh"
#include "../../efficientlo-net_b200/csrc/elo_tc.cuh"
#include "elo_b200_test.h"

namespace elo {

__global__ void __launch_bounds__(256, 1) tc_dense_test_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                               const float* __restrict__ bias, float* __restrict__ Y,
                                                               int K, int N, int relu)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Bhi = reinterpret_cast<float*>(smem_raw);
    float* Blo = Bhi + (size_t)K * N;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::TMEM_COLS);
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    // B operand: [K/4][N][4], split into tf32 hi / lo parts
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i - k * N;
        uint32_t hi, lo;
        tc::split_tf32(__ldg(W + i), hi, lo);
        const int dst = ((k >> 2) * N + n) * 4 + (k & 3);
        Bhi[dst] = __uint_as_float(hi);
        Blo[dst] = __uint_as_float(lo);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int m = 32 * (warp & 3) + lane;
    const uint32_t colA_hi = 0, colA_lo = (uint32_t)K, colD = 384;

    // A operand: this thread's row, 16 columns at a time
    for (int cb = warp >> 2; cb < K / 16; cb += 2) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) tc::split_tf32(__ldg(X + (size_t)m * K + cb * 16 + i), hi[i], lo[i]);
        tc::tmem_st16(tbase + lane_base + colA_hi + cb * 16, hi);
        tc::tmem_st16(tbase + lane_base + colA_lo + cb * 16, lo);
    }
    tc::tmem_st_wait();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem written by threads, read by the MMA
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_tf32((uint32_t)N);
        const uint32_t lbo = (uint32_t)N * 16, sbo = 128;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t bhi = tc::smem_desc(tc::smem_addr(Bhi + (size_t)ks * 2 * N * 4), lbo, sbo);
            const uint64_t blo = tc::smem_desc(tc::smem_addr(Blo + (size_t)ks * 2 * N * 4), lbo, sbo);
            tc::mma_ts(tbase + colD, tbase + colA_hi + ks * 8, bhi, idesc, ks > 0);
            tc::mma_ts(tbase + colD, tbase + colA_lo + ks * 8, bhi, idesc, true);
            tc::mma_ts(tbase + colD, tbase + colA_hi + ks * 8, blo, idesc, true);
        }
        tc::mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc::fence_after_sync();
    for (int cb = warp >> 2; cb < N / 16; cb += 2) {
        float v[16];
        tc::tmem_ld16(tbase + lane_base + colD + cb * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float y = v[i] + __ldg(bias + cb * 16 + i);
            if (relu) y = fmaxf(y, 0.f);
            Y[(size_t)m * N + cb * 16 + i] = y;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::TMEM_COLS);
}

// Throughput probe: `iters` back-to-back MMAs (128 x N x 8, tf32) into one accumulator, A from TMEM (ts = 1) or
// from shared memory (ts = 0); out[0] = SM cycles from first issue to completion.
__global__ void __launch_bounds__(128, 1) tc_mma_bench_kernel(int N, int iters, int ts, int nacc, long long* out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* B = reinterpret_cast<float*>(smem_raw);            // [2][N][4]
    float* A = B + 2 * 128 * 4;                               // [2][128][4] (SS form)
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, tc::TMEM_COLS);
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < 4 * 128 * 4; i += blockDim.x) B[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    {   // zero the A columns in TMEM
        uint32_t z[16];
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        tc::tmem_st16(tbase + ((uint32_t)(32 * (warp & 3)) << 16), z);
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_tf32((uint32_t)N);
        const uint64_t bdesc = tc::smem_desc(tc::smem_addr(B), (uint32_t)N * 16, 128);
        const uint64_t adesc = tc::smem_desc(tc::smem_addr(A), 128 * 16, 128);
        const long long t0 = clock64();
        const uint32_t d0 = tbase + 128, d1 = d0 + (nacc > 1 ? N : 0), d2 = d0 + (nacc > 2 ? 2 * N : 0);
        (void)adesc;
        for (int i = 0; i < iters; i += 3) {          // three MMAs per trip, rotating over the accumulators
            tc::mma_ts_c<true>(d0, tbase, bdesc, idesc);
            tc::mma_ts_c<true>(d1, tbase + (ts ? 8 : 0), bdesc, idesc);
            tc::mma_ts_c<true>(d2, tbase, bdesc, idesc);
        }
        tc::mma_commit(&mbar);
        mbar_wait(&mbar, 0);
        out[0] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, tc::TMEM_COLS);
}

}  // namespace elo

extern "C" int elo_tc_mma_bench(int N, int iters, int ts, int nacc, long long* out_cycles, void* stream)
{
    using namespace elo;
    if (nacc < 1 || nacc * N > 384) return set_error(ELO_ERR_INVALID_ARGUMENT, "tc_mma_bench: nacc * N <= 384");
    tc_mma_bench_kernel<<<1, 128, 16384, (cudaStream_t)stream>>>(N, iters, ts, nacc, out_cycles);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "tc_mma_bench launch");
}

extern "C" int elo_tc_dense_test(const float* X, const float* W, const float* bias, float* Y, int K, int N, int relu,
                                 void* stream)
{
    using namespace elo;
    if (!X || !W || !bias || !Y || K <= 0 || K > 192 || (K & 15) || (N != 64 && N != 128))
        return set_error(ELO_ERR_INVALID_ARGUMENT, "tc_dense_test: K % 16 == 0, K <= 192, N in {64,128}");
    const size_t smem = (size_t)2 * K * N * 4;
    cudaError_t err = cudaFuncSetAttribute(tc_dense_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return set_cuda_error(err, "tc_dense_test smem");
    tc_dense_test_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(X, W, bias, Y, K, N, relu);
    count_launches(1);
    err = cudaGetLastError();
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "tc_dense_test launch");
}
