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
    uint32_t* tmem_holder;    // shared-memory word tcgen05.alloc writes the base address to
    uint32_t layer;       // layers completed so far (parity of a_ready / done)
    uint32_t chunk;       // chunks consumed so far (MMA warp)
    uint32_t cslot, cround;   // ring slot / lap of the next chunk (MMA warp)
    const float* bias;    // biases of the remaining layers, in shared memory (compute warps advance it)
    long long* tlog;      // optional phase timestamps (ns) of CTA (0,0): compute thread 0 -> [0,32), MMA thread -> [32,64)

    __device__ __forceinline__ void stamp(int id)
    {
        if (tlog != nullptr && blockIdx.x == 0 && blockIdx.y == 0 &&
            (threadIdx.x == 0 || (threadIdx.x >> 5) == COMPUTE_WARPS + 1)) {
            long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            tlog[id + ((threadIdx.x >> 5) == 0 ? 0 : 32)] = t;
        }
    }

    // all threads; bars must hold 2 * MAX_RING + 2 mbarriers; contains __syncthreads
    // Set-up, split in two so that a kernel can put its first loads of upstream data between the halves:
    //   begin(): per-role work on constants only -- the TMA warp initialises the mbarriers, the MMA warp copies
    //            the biases to shared memory, warp 0 allocates tensor memory;
    //   join():  CTA-wide barrier that publishes all of it.
    // Under programmatic dependent launch begin() runs while the previous kernel is still finishing; the compute
    // warps call pdl_wait() after it, the other two warps never touch upstream data and do not wait at all.
    // alloc_now = false: tensor memory is claimed later, by alloc_late(), right before the first A operand is
    // written -- with a second CTA resident on the SM (the weight ring is sized for that) one tile's set-up and
    // gather then overlap the other's MMA chain, and finish() can be called before the pooling for the same reason.
    __device__ __forceinline__ void begin(float* ring_, uint64_t* bars, uint32_t nring_, uint32_t* tmem_holder_,
                                          const float* bias_gmem, float* bias_smem, int nbias, long long* tlog_ = nullptr,
                                          bool alloc_now = true)
    {
        tlog = tlog_;
        stamp(0);
        ring = ring_; full = bars; empty = bars + MAX_RING; a_ready = bars + 2 * MAX_RING; done = a_ready + 1;
        nring = nring_; layer = 0; chunk = 0; cslot = 0; cround = 0; bias = bias_smem;
        tmem_holder = tmem_holder_;
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp == COMPUTE_WARPS && lane == 0) {
            for (uint32_t i = 0; i < nring; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
            mbar_init(a_ready, COMPUTE_WARPS);
            mbar_init(done, 1);
            mbar_fence_init();
        }
        // every epilogue needs its biases at once: one L2 round trip here instead of one per layer
        if (warp == COMPUTE_WARPS + 1)
            for (int i = lane; i < nbias; i += 32) bias_smem[i] = __ldg(bias_gmem + i);
        if (warp == 0 && alloc_now) tc::tmem_alloc(tmem_holder, tc::TMEM_COLS);
    }
    // compute warps only (all of them): blocks until the SM's tensor memory is free
    __device__ __forceinline__ void alloc_late()
    {
        if ((threadIdx.x >> 5) == 0) tc::tmem_alloc(tmem_holder, tc::TMEM_COLS);
        tc::fence_before_sync();
        compute_sync();
        tc::fence_after_sync();
        tbase = *reinterpret_cast<volatile uint32_t*>(tmem_holder);
    }
    __device__ __forceinline__ void join()
    {
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        tbase = *tmem_holder;
        stamp(1);
    }
    __device__ __forceinline__ void init(float* ring_, uint64_t* bars, uint32_t nring_, uint32_t* tmem_holder_,
                                         const float* bias_gmem, float* bias_smem, int nbias, long long* tlog_ = nullptr)
    {
        begin(ring_, bars, nring_, tmem_holder_, bias_gmem, bias_smem, nbias, tlog_);
        join();
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
        if (!tc::elect_one()) return;
        uint32_t slot = 0, round = 0;
        for (uint32_t g = 0; g < total; ++g) {
            if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
            mbar_expect_tx(full + slot, TC_CHUNK_BYTES);
            bulk_g2s(ring + (size_t)slot * TC_CHUNK_FLOATS, src + (size_t)g * TC_CHUNK_FLOATS, TC_CHUNK_BYTES, full + slot);
            if (++slot == nring) { slot = 0; ++round; }
        }
    }

    // ---- MMA warp (lane 0) ------------------------------------------------------------------------
    // D[128 x N] = A[:, a_col : a_col + cin] * W  with the 3xTF32 split.  The issuing thread is the
    // serial bottleneck of a small-N layer (a 128 x 64 x 8 MMA occupies the tensor pipe for ~35 cycles),
    // so the loop is kept to a few instructions per MMA: descriptors are advanced by adds, N is a
    // compile-time constant and the accumulate flags are immediates.
    template <uint32_t N>
    __device__ __forceinline__ void issue_layer_n(uint32_t a_col, uint32_t cin)
    {
        constexpr uint32_t R = 2048u / N;              // k-rows per chunk
        constexpr uint32_t KSPC = R / 8u;              // K=8 steps per chunk
        constexpr uint32_t KSTEP16 = 2u * N;           // 16-byte units between consecutive K-steps of B
        mbar_wait(a_ready, layer & 1u);
        tc::fence_after_sync();
        tbase = *reinterpret_cast<volatile uint32_t*>(tmem_holder);    // late allocation: known once A is ready
        stamp(2 + 2 * (int)layer);
        const uint32_t nks = (cin + 7u) >> 3;
        const uint32_t idesc = tc::idesc_tf32(N);
        const uint64_t desc0 = tc::smem_desc(tc::smem_addr(ring), N * 16u, 128u);
        const uint32_t d = tbase + TC_D_COL;
        uint32_t a_hi = tbase + a_col;
        uint32_t ks = 0;
        while (ks < nks) {
            const uint32_t slot = cslot;
            mbar_wait(full + slot, cround & 1u);
            tc::fence_after_sync();
            uint64_t bhi = desc0 + (uint64_t)slot * (TC_CHUNK_BYTES >> 4);
            uint64_t blo = bhi + (TC_CHUNK_BYTES >> 5);
#pragma unroll
            for (uint32_t j = 0; j < KSPC; ++j) {
                if (ks < nks) {
                    if (ks == 0) tc::mma_ts_c<false>(d, a_hi, bhi, idesc);
                    else tc::mma_ts_c<true>(d, a_hi, bhi, idesc);
                    tc::mma_ts_c<true>(d, a_hi + TC_A_LO, bhi, idesc);
                    tc::mma_ts_c<true>(d, a_hi, blo, idesc);
                    a_hi += 8u; bhi += KSTEP16; blo += KSTEP16; ++ks;
                }
            }
            tc::mma_commit(empty + slot);              // slot reusable once these MMAs have read it
            ++chunk;
            if (++cslot == nring) { cslot = 0; ++cround; }
        }
        tc::mma_commit(done);
        stamp(3 + 2 * (int)layer);
        ++layer;
    }
    __device__ __forceinline__ void issue_layer(uint32_t a_col, uint32_t cin, uint32_t n)
    {
        if (n == 128u) issue_layer_n<128u>(a_col, cin);
        else issue_layer_n<64u>(a_col, cin);
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
        stamp(8 + 2 * (int)layer);
        const uint32_t addr = my_lane_addr() + TC_D_COL;
        for (int b = my_half(); b * 16 < n; b += 2) {
            float v[16];
            tc::tmem_ld16(addr + b * 16, v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(bias + b * 16 + i);
                v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
            }
            if (RELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            sink(b, v);
        }
        bias += n;
        stamp(9 + 2 * (int)layer);
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
