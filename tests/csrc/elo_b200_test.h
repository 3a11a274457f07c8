/* elo_b200_test.h -- TEST HOOKS, built into tests/csrc/libelo_b200_test.so (never into the product library):
 * direct doors onto the tcgen05 primitives of efficientlo-net_b200/csrc/elo_tc.cuh. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

/* Test hook for the tensor-core dense layer (tcgen05, 3xTF32): Y[128 x N] = act(X[128 x K] W[K x N] + bias),
 * X, W, Y row-major on the device, K % 16 == 0, K <= 192, N in {64, 128}. */
int elo_tc_dense_test(const float *X, const float *W, const float *bias, float *Y, int K, int N, int relu,
                      void *stream);

/* Throughput probe: `iters` tf32 MMAs (128 x N x 8) rotating over `nacc` accumulators; out_cycles[0] (device) = SM cycles. */
int elo_tc_mma_bench(int N, int iters, int ts, int nacc, long long *out_cycles, void *stream);

#ifdef __cplusplus
}
#endif
