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
//     shared by every CTA), cut into 8 KB chunks; a single thread moves chunk after chunk into a
//     ring in shared memory with cp.async.bulk (the TMA engine, SASS UBLKCP) completing on
//     mbarriers, running ahead of the math by RING-1 chunks;
//   * 256 threads as 16 x 16: a thread owns 4*NB rows x (COUT/16) columns in registers
//     (FFMA; fp32 end to end, which is what holds the 1e-4 parity bar -- DESIGN.md section 5).
//
// Packed layer format (host: efficientlo-net_b200/packing.py): row 0 = folded bias, rows 1..cin =
// W'[k][0..COUT), zero rows up to a multiple of CHUNK_FLOATS / COUT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elo {

constexpr int CTA_THREADS = 256;
constexpr int CHUNK_FLOATS = 2048;                 // 8 KB per weight chunk
constexpr int CHUNK_BYTES = CHUNK_FLOATS * 4;
constexpr int RING = 4;                            // chunks in flight
constexpr int RING_BYTES = RING * CHUNK_BYTES;

// ---- shared-memory activation layout -----------------------------------------------------------
// element (channel c, row r) of a buffer with RS rows per channel
__device__ __forceinline__ int act_index(int c, int r, int RS)
{
    return c * RS + ((((r >> 2) ^ ((c >> 2) & 7))) << 2) + (r & 3);
}

// ---- mbarrier / bulk-copy primitives (PTX) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy through the TMA engine; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- weight stream -----------------------------------------------------------------------------
// All threads call acquire()/release() in lock step; thread 0 is the producer.
struct WeightStream {
    const float* src;      // packed weights of the kernel (global, 16 B aligned)
    float* ring;           // RING x CHUNK_FLOATS (shared, 16 B aligned)
    uint64_t* full;        // RING mbarriers (shared)
    uint32_t total;        // chunks in `src`; the sequence wraps (persistent CTAs re-read it per tile)
    uint32_t consumed;     // chunks consumed so far by this CTA
    uint32_t issued;       // chunks issued so far (meaningful in thread 0)
    uint32_t limit;        // total chunks this CTA will ever consume (tiles * total)

    __device__ __forceinline__ void issue_one()
    {
        const uint32_t slot = issued % RING;
        const uint32_t chunk = issued % total;
        mbar_expect_tx(full + slot, CHUNK_BYTES);
        bulk_g2s(ring + (size_t)slot * CHUNK_FLOATS, src + (size_t)chunk * CHUNK_FLOATS, CHUNK_BYTES, full + slot);
        ++issued;
    }
    // call once by all threads, before the first acquire (contains a __syncthreads)
    __device__ __forceinline__ void start(const float* src_, float* ring_, uint64_t* full_, uint32_t total_,
                                          uint32_t tiles)
    {
        src = src_; ring = ring_; full = full_; total = total_;
        consumed = 0; issued = 0; limit = total_ * tiles;
        if (threadIdx.x == 0) {
            for (int i = 0; i < RING; ++i) mbar_init(full + i, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            while (issued < (uint32_t)RING && issued < limit) issue_one();
    }
    __device__ __forceinline__ const float* acquire()
    {
        const uint32_t slot = consumed % RING;
        mbar_wait(full + slot, (consumed / RING) & 1u);
        return ring + (size_t)slot * CHUNK_FLOATS;
    }
    // every thread is done reading the current chunk -> its slot can be refilled
    __device__ __forceinline__ void release()
    {
        __syncthreads();
        ++consumed;
        if (threadIdx.x == 0 && issued < limit) issue_one();
    }
};

// ---- one dense layer on the tile -----------------------------------------------------------------
// out[n][row] = act( bias[n] + sum_k in[k][row] * W[k][n] ),  k < cin,  n < COUT,  row < 64*NB.
// `in` / `out` must start at a channel that is a multiple of 32 of their buffer (swizzle phase).
// Ends with a __syncthreads(): `out` is visible to all threads on return.
template <int NB, int COUT, bool RELU>
__device__ __forceinline__ void dense(WeightStream& ws, const float* __restrict__ in, int cin,
                                      float* __restrict__ out)
{
    constexpr int NV = COUT / 64;              // float4 column groups per thread
    constexpr int R = CHUNK_FLOATS / COUT;     // weight rows per chunk
    constexpr int RS = NB * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[NB][4][NV * 4];
    const int nrows = cin + 1;
    const int nchunks = (nrows + R - 1) / R;
    for (int ch = 0; ch < nchunks; ++ch) {
        const float* w = ws.acquire();
        int r = 0;
        if (ch == 0) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 b = *reinterpret_cast<const float4*>(w + v * 64 + tx * 4);
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[nb][i][v * 4 + 0] = b.x; acc[nb][i][v * 4 + 1] = b.y;
                        acc[nb][i][v * 4 + 2] = b.z; acc[nb][i][v * 4 + 3] = b.w;
                    }
            }
            r = 1;
        }
        const int r_end = min(R, nrows - ch * R);
        const int kbase = ch * R - 1;
#pragma unroll 4
        for (; r < r_end; ++r) {
            const int k = kbase + r;
            float4 wv[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) wv[v] = *reinterpret_cast<const float4*>(w + r * COUT + v * 64 + tx * 4);
            const float* arow = in + k * RS + ((ty ^ ((k >> 2) & 7)) << 2);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const float4 a = *reinterpret_cast<const float4*>(arow + nb * 64);
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        acc[nb][i][v * 4 + 0] = fmaf(av[i], wv[v].x, acc[nb][i][v * 4 + 0]);
                        acc[nb][i][v * 4 + 1] = fmaf(av[i], wv[v].y, acc[nb][i][v * 4 + 1]);
                        acc[nb][i][v * 4 + 2] = fmaf(av[i], wv[v].z, acc[nb][i][v * 4 + 2]);
                        acc[nb][i][v * 4 + 3] = fmaf(av[i], wv[v].w, acc[nb][i][v * 4 + 3]);
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
    __syncthreads();
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
