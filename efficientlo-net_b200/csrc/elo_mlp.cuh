// elo_mlp.cuh -- per-group shared MLP (chained 1x1 convs with folded BN + ReLU) on one CTA tile.
//
// What the reference does with ~9 TensorFlow kernels per layer (cuDNN 1x1 conv, bias, FusedBatchNorm,
// ReLU, each a round trip of the (B,N,K,C) tensor through HBM -- utils/tf_util.py:120-185 as called
// from utils/pointnet_util.py:72-90, 131-135, 217-222, 289-311) happens here on a tile of rows that
// never leaves shared memory:
//
//   * activations live in shared memory as [channel][row] (rows fastest), 64*NB rows per tile, with
//     the 4-row quads XOR-swizzled by the channel so that both the k-loop's float4 loads and the
//     epilogue's float4 stores are bank-conflict free;
//   * the weights of all layers of a kernel are ONE packed stream in global memory (L2 resident,
//     shared by every CTA), cut into 8 KB chunks; a dedicated producer warp moves chunk after chunk
//     into a ring in shared memory with cp.async.bulk (the TMA engine, SASS UBLKCP) completing on
//     "full" mbarriers, and refills a slot as soon as all eight compute warps have arrived on its
//     "empty" mbarrier -- the math never meets a CTA-wide barrier inside a layer;
//   * 256 threads as 16 x 16: a thread owns 4*NB rows x (COUT/16) columns in registers
//     (FFMA; fp32 end to end, which is what holds the 1e-4 parity bar -- DESIGN.md section 5).
//
// Packed layer format (host: efficientlo-net_b200/packing.py): rows 0..cin-1 = W'[k][0..COUT), row cin =
// folded bias, zero rows up to a multiple of CHUNK_FLOATS / COUT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "elo_bulk.cuh"

namespace elo {

constexpr int CTA_THREADS = 256;                   // compute threads (8 warps)
constexpr int LAUNCH_THREADS = CTA_THREADS + 32;   // + one producer warp that only drives the TMA engine
constexpr int COMPUTE_WARPS = CTA_THREADS / 32;
constexpr int CHUNK_FLOATS = 2048;                 // 8 KB per weight chunk
constexpr int CHUNK_BYTES = CHUNK_FLOATS * 4;
constexpr int MAX_RING = 8;                        // chunks in flight (host picks 2..8 by the shared memory left)

// ---- shared-memory activation layout -----------------------------------------------------------
// element (channel c, row r) of a buffer with RS rows per channel
__device__ __forceinline__ int act_index(int c, int r, int RS)
{
    return c * RS + ((((r >> 2) ^ ((c >> 2) & 7))) << 2) + (r & 3);
}

// barrier over the 256 compute threads only (the producer warp never joins it)
__device__ __forceinline__ void compute_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(CTA_THREADS) : "memory");
}

// ---- weight stream -----------------------------------------------------------------------------
// Producer / consumer ring of weight chunks in shared memory:
//   full[s]  (count 1)             armed by the producer with expect_tx, completed by the TMA bytes
//   empty[s] (count COMPUTE_WARPS) one arrival per compute warp when it is done reading slot s
// The producer warp (warp COMPUTE_WARPS) issues chunk after chunk, blocking only when the ring is full;
// compute warps block only when their next chunk has not landed.  No CTA-wide barrier per chunk.
struct WeightStream {
    float* ring;           // nring x CHUNK_FLOATS (shared, 16 B aligned)
    uint64_t* full;        // nring mbarriers
    uint64_t* empty;       // nring mbarriers
    uint32_t nring;
    uint32_t consumed;     // chunks consumed so far by this warp

    // all threads, once; contains a CTA-wide __syncthreads (producer warp included)
    __device__ __forceinline__ void init(float* ring_, uint64_t* bars, uint32_t nring_)
    {
        ring = ring_; full = bars; empty = bars + MAX_RING; nring = nring_; consumed = 0;
        if (threadIdx.x == 0) {
            for (uint32_t i = 0; i < nring; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, COMPUTE_WARPS); }
            mbar_fence_init();
        }
        __syncthreads();
    }
    // producer warp: stream `total` chunks of `src`, `passes` times over
    __device__ __forceinline__ void produce(const float* src, uint32_t total, uint32_t passes)
    {
        if ((threadIdx.x & 31) != 0) return;
        const uint32_t limit = total * passes;
        for (uint32_t g = 0; g < limit; ++g) {
            const uint32_t slot = g % nring, round = g / nring;
            if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
            mbar_expect_tx(full + slot, CHUNK_BYTES);
            bulk_g2s(ring + (size_t)slot * CHUNK_FLOATS, src + (size_t)(g % total) * CHUNK_FLOATS, CHUNK_BYTES, full + slot);
        }
    }
    __device__ __forceinline__ const float* acquire()
    {
        const uint32_t slot = consumed % nring;
        mbar_wait(full + slot, (consumed / nring) & 1u);
        return ring + (size_t)slot * CHUNK_FLOATS;
    }
    // this warp is done reading the current chunk
    __device__ __forceinline__ void release()
    {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + consumed % nring);
        ++consumed;
    }
};

// ---- one dense layer on the tile -----------------------------------------------------------------
// out[n][row] = act( bias[n] + sum_k in[k][row] * W[k][n] ),  k < cin,  n < COUT,  row < 64*NB.
// `in` / `out` must start at a channel that is a multiple of 32 of their buffer (swizzle phase).
// Called by the 256 compute threads.  Ends with compute_sync(): `out` is visible to all of them on return.
template <int NB, int COUT, bool RELU>
__device__ __forceinline__ void dense(WeightStream& ws, const float* __restrict__ in, int cin,
                                      float* __restrict__ out)
{
    constexpr int NV = COUT / 64;              // float4 column groups per thread
    constexpr int R = CHUNK_FLOATS / COUT;     // weight rows per chunk (16 or 32: a multiple of 4)
    constexpr int RS = NB * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[NB][4][NV * 4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < NV * 4; ++c) acc[nb][i][c] = 0.f;

    auto fma_step = [&](const float4 (&wv)[NV], const float4 (&av)[NB]) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const float a4[4] = {av[nb].x, av[nb].y, av[nb].z, av[nb].w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc[nb][i][v * 4 + 0] = fmaf(a4[i], wv[v].x, acc[nb][i][v * 4 + 0]);
                    acc[nb][i][v * 4 + 1] = fmaf(a4[i], wv[v].y, acc[nb][i][v * 4 + 1]);
                    acc[nb][i][v * 4 + 2] = fmaf(a4[i], wv[v].z, acc[nb][i][v * 4 + 2]);
                    acc[nb][i][v * 4 + 3] = fmaf(a4[i], wv[v].w, acc[nb][i][v * 4 + 3]);
                }
        }
    };

    const int nchunks = (cin + 1 + R - 1) / R;          // rows 0..cin-1 = W[k][:], row cin = bias
    for (int ch = 0; ch < nchunks; ++ch) {
        const float* w = ws.acquire() + tx * 4;
        const int kcount = min(R, cin - ch * R);         // weight rows in this chunk (<= 0: bias only)
        const int groups = kcount > 0 ? kcount >> 2 : 0;
        // four k-steps at a time: k0 is a multiple of 4, so the swizzle phase (k >> 2) & 7 is shared and
        // every address below is a compile-time offset from two base pointers
        for (int g4 = 0; g4 < groups; ++g4) {
            const int k0 = ch * R + g4 * 4;
            const float* a0 = in + k0 * RS + ((ty ^ ((k0 >> 2) & 7)) << 2);
            const float* w0 = w + g4 * 4 * COUT;
            float4 wv[4][NV], av[4][NB];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int v = 0; v < NV; ++v) wv[j][v] = *reinterpret_cast<const float4*>(w0 + j * COUT + v * 64);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) av[j][nb] = *reinterpret_cast<const float4*>(a0 + j * RS + nb * 64);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) fma_step(wv[j], av[j]);
        }
        for (int r = groups * 4; r < kcount; ++r) {      // tail of a layer whose cin is not a multiple of 4
            const int k = ch * R + r;
            float4 wv[NV], av[NB];
#pragma unroll
            for (int v = 0; v < NV; ++v) wv[v] = *reinterpret_cast<const float4*>(w + r * COUT + v * 64);
            const float* a0 = in + k * RS + ((ty ^ ((k >> 2) & 7)) << 2);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) av[nb] = *reinterpret_cast<const float4*>(a0 + nb * 64);
            fma_step(wv, av);
        }
        if (ch == nchunks - 1) {                         // the bias row lives in the last chunk
            const float* brow = w + (cin - ch * R) * COUT;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 b = *reinterpret_cast<const float4*>(brow + v * 64);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[nb][i][v * 4 + 0] += b.x; acc[nb][i][v * 4 + 1] += b.y;
                        acc[nb][i][v * 4 + 2] += b.z; acc[nb][i][v * 4 + 3] += b.w;
                    }
            }
        }
        ws.release();
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = v * 64 + tx * 4 + j;
            float* orow = out + n * RS + ((ty ^ ((n >> 2) & 7)) << 2);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                float4 o = make_float4(acc[nb][0][v * 4 + j], acc[nb][1][v * 4 + j], acc[nb][2][v * 4 + j],
                                       acc[nb][3][v * 4 + j]);
                if (RELU) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                *reinterpret_cast<float4*>(orow + nb * 64) = o;
            }
        }
    compute_sync();
}

// runtime width dispatch (every layer on this path is 64 or 128 wide)
template <int NB>
__device__ __forceinline__ void dense_rt(WeightStream& ws, const float* in, int cin, float* out, int cout)
{
    if (cout == 128) dense<NB, 128, true>(ws, in, cin, out);
    else dense<NB, 64, true>(ws, in, cin, out);
}

// chunks a layer occupies in the packed stream
__host__ __device__ inline int layer_chunks(int cin, int cout)
{
    const int R = CHUNK_FLOATS / cout;
    return (cin + 1 + R - 1) / R;
}

}  // namespace elo
