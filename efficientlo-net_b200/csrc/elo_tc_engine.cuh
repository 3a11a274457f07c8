// elo_tc_engine.cuh -- the 128-row tensor-core tile engine behind the fused MLP kernels (sm_100a).
//
// Roles inside one CTA of TC_LAUNCH_THREADS threads:
//   warps 0..7  compute   search / gather into shared memory, TMEM<->register epilogues, pooling
//   warp  8     TMA       streams the packed weights (16 KB chunks) into a shared-memory ring
//   warp  9     MMA       one lane issues tcgen05.mma for every layer
// Synchronisation is by mbarriers only:
//   full[s]/empty[s]   weight ring (TMA completes full; tcgen05.commit releases empty)
//   a_ready            the 8 compute warps -> MMA warp: "the A operand of the next layer is in TMEM"
//   done               tcgen05.commit -> compute warps: "the accumulator of this layer is complete"
//
// TMEM map (512 columns x 128 lanes, lane = row of the tile):
//   [  0,192)  A operand, tf32 hi parts     [192,384)  A operand, lo parts      [384,512)  accumulator D
// A compute thread owns row 32*(warp&3)+lane (the lane quarter tcgen05.ld/st lets its warp touch) and
// the 16-column blocks b with b % 2 == warp >> 2.
//
// Packed TC weight stream (host: packing.pack_stream_tc): per layer, ceil(cin / R) chunks with
// R = 2048 / N k-rows each; a chunk is [hi: (R/4) x N x 4 floats][lo: same] in the canonical K-major
// core-matrix order; k-rows beyond cin are zero.  Biases travel separately (one float per channel).
#pragma once
#include "elo_mlp.cuh"
#include "elo_tc.cuh"

namespace elo {

constexpr int TC_ROWS = 128;
constexpr int TC_LAUNCH_THREADS = CTA_THREADS + 64;
constexpr int TC_CHUNK_FLOATS = 4096;
constexpr int TC_CHUNK_BYTES = TC_CHUNK_FLOATS * 4;
constexpr uint32_t TC_A_LO = 192;     // column offset of the lo parts
constexpr uint32_t TC_D_COL = 384;    // accumulator columns

struct TcPipe {
    float* ring;
    uint64_t* full;       // [MAX_RING]
    uint64_t* empty;      // [MAX_RING]
    uint64_t* a_ready;
    uint64_t* done;
    uint32_t nring;
    uint32_t tbase;       // TMEM base address
    uint32_t layer;       // layers completed so far (parity of a_ready / done)
    uint32_t chunk;       // chunks consumed so far (MMA warp)
    const float* bias;    // biases of the remaining layers (compute warps advance it)

    // all threads; bars must hold 2 * MAX_RING + 2 mbarriers; contains __syncthreads
    __device__ __forceinline__ void init(float* ring_, uint64_t* bars, uint32_t nring_, uint32_t* tmem_holder,
                                         const float* bias_)
    {
        ring = ring_; full = bars; empty = bars + MAX_RING; a_ready = bars + 2 * MAX_RING; done = a_ready + 1;
        nring = nring_; layer = 0; chunk = 0; bias = bias_;
        if (threadIdx.x == 0) {
            for (uint32_t i = 0; i < nring; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
            mbar_init(a_ready, COMPUTE_WARPS);
            mbar_init(done, 1);
            mbar_fence_init();
        }
        if ((threadIdx.x >> 5) == 0) tc::tmem_alloc(tmem_holder, tc::TMEM_COLS);
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        tbase = *tmem_holder;
    }
    // after the last epilogue: compute warps only
    __device__ __forceinline__ void finish()
    {
        tc::fence_before_sync();
        compute_sync();
        if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(tbase, tc::TMEM_COLS);
    }

    // ---- TMA warp ------------------------------------------------------------------------------
    __device__ __forceinline__ void produce(const float* src, uint32_t total)
    {
        if ((threadIdx.x & 31) != 0) return;
        for (uint32_t g = 0; g < total; ++g) {
            const uint32_t slot = g % nring, round = g / nring;
            if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
            mbar_expect_tx(full + slot, TC_CHUNK_BYTES);
            bulk_g2s(ring + (size_t)slot * TC_CHUNK_FLOATS, src + (size_t)g * TC_CHUNK_FLOATS, TC_CHUNK_BYTES, full + slot);
        }
    }

    // ---- MMA warp (lane 0) ------------------------------------------------------------------------
    // D[128 x n] = A[:, a_col : a_col + cin] * W  with the 3xTF32 split
    __device__ __forceinline__ void issue_layer(uint32_t a_col, uint32_t cin, uint32_t n)
    {
        mbar_wait(a_ready, layer & 1u);
        tc::fence_after_sync();
        const uint32_t R = 2048u / n;                  // k-rows per chunk
        const uint32_t nks = (cin + 7u) >> 3;          // K = 8 per MMA
        const uint32_t nchunks = (cin + R - 1u) / R;
        const uint32_t idesc = tc::idesc_tf32(n);
        const uint32_t lbo = n * 16u, sbo = 128u;
        uint32_t ks = 0;
        for (uint32_t c = 0; c < nchunks; ++c) {
            const uint32_t slot = chunk % nring;
            mbar_wait(full + slot, (chunk / nring) & 1u);
            tc::fence_after_sync();
            const float* hi = ring + (size_t)slot * TC_CHUNK_FLOATS;
            const float* lo = hi + TC_CHUNK_FLOATS / 2;
            for (uint32_t j = 0; j < R / 8u && ks < nks; ++j, ++ks) {
                const uint64_t bhi = tc::smem_desc(tc::smem_addr(hi + (size_t)j * 2u * n * 4u), lbo, sbo);
                const uint64_t blo = tc::smem_desc(tc::smem_addr(lo + (size_t)j * 2u * n * 4u), lbo, sbo);
                const uint32_t a_hi = tbase + a_col + ks * 8u, a_lo = a_hi + TC_A_LO;
                tc::mma_ts(tbase + TC_D_COL, a_hi, bhi, idesc, ks > 0);
                tc::mma_ts(tbase + TC_D_COL, a_lo, bhi, idesc, true);
                tc::mma_ts(tbase + TC_D_COL, a_hi, blo, idesc, true);
            }
            tc::mma_commit(empty + slot);              // slot reusable once these MMAs have read it
            ++chunk;
        }
        tc::mma_commit(done);
        ++layer;
    }

    // ---- compute warps ------------------------------------------------------------------------------
    __device__ __forceinline__ uint32_t my_lane_addr() const { return tbase + ((uint32_t)(32 * ((threadIdx.x >> 5) & 3)) << 16); }
    __device__ __forceinline__ int my_row() const { return 32 * ((threadIdx.x >> 5) & 3) + (threadIdx.x & 31); }
    __device__ __forceinline__ int my_half() const { return (threadIdx.x >> 5) >> 2; }

    // TMEM A[:, a_col + i] = X[c0 + i][row] for i < ncols (X in the swizzled [channel][row] layout, RS = 128),
    // zero up to the next multiple of 16 columns.  a_col must be a multiple of 16.
    __device__ __forceinline__ void load_a_from_smem(const float* X, int c0, int ncols, uint32_t a_col)
    {
        const int m = my_row();
        const uint32_t addr = my_lane_addr() + a_col;
        for (int b = my_half(); b * 16 < ncols; b += 2) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = b * 16 + i;
                const float v = c < ncols ? X[act_index(c0 + c, m, TC_ROWS)] : 0.f;
                tc::split_tf32(v, hi[i], lo[i]);
            }
            tc::tmem_st16(addr + b * 16, hi);
            tc::tmem_st16(addr + TC_A_LO + b * 16, lo);
        }
    }
    // same, from per-thread values produced by a functor val(column) (used to re-materialise small inputs)
    template <typename F>
    __device__ __forceinline__ void load_a_from_regs(int ncols, uint32_t a_col, F val)
    {
        const uint32_t addr = my_lane_addr() + a_col;
        for (int b = my_half(); b * 16 < ncols; b += 2) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = b * 16 + i;
                tc::split_tf32(c < ncols ? val(c) : 0.f, hi[i], lo[i]);
            }
            tc::tmem_st16(addr + b * 16, hi);
            tc::tmem_st16(addr + TC_A_LO + b * 16, lo);
        }
    }
    // the A operand of the next layer is complete (call by all compute threads)
    __device__ __forceinline__ void signal_a_ready()
    {
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(a_ready);
    }
    // Wait for the current layer's accumulator, then for each of this thread's 16-column blocks apply
    // bias (+ ReLU) and hand the 16 values to sink(block, values[16]).  Advances to the next layer.
    template <bool RELU, typename Sink>
    __device__ __forceinline__ void epilogue(int n, Sink sink)
    {
        mbar_wait(done, layer & 1u);
        tc::fence_after_sync();
        const uint32_t addr = my_lane_addr() + TC_D_COL;
        for (int b = my_half(); b * 16 < n; b += 2) {
            float v[16];
            tc::tmem_ld16(addr + b * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] += __ldg(bias + b * 16 + i);
                if (RELU) v[i] = fmaxf(v[i], 0.f);
            }
            sink(b, v);
        }
        bias += n;
        ++layer;
    }
    // sink helpers
    __device__ __forceinline__ void store_a(uint32_t a_col, int b, const float (&v)[16])
    {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) tc::split_tf32(v[i], hi[i], lo[i]);
        const uint32_t addr = my_lane_addr() + a_col + b * 16;
        tc::tmem_st16(addr, hi);
        tc::tmem_st16(addr + TC_A_LO, lo);
    }
    __device__ __forceinline__ void store_smem(float* S, int c0, int b, const float (&v)[16])
    {
        const int m = my_row();
#pragma unroll
        for (int i = 0; i < 16; ++i) S[act_index(c0 + b * 16 + i, m, TC_ROWS)] = v[i];
    }
};

// chunks a layer occupies in the packed TC stream
__host__ __device__ inline int tc_layer_chunks(int cin, int n)
{
    const int R = 2048 / n;
    return (cin + R - 1) / R;
}

}  // namespace elo
