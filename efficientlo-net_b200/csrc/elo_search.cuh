// elo_search.cuh -- projection-aware neighbour search, one warp per query (sm_100a).
//
// Device-side core shared by the stand-alone index ops (fused_conv_index.cu, the drop-in for
// the reference's FusedConv{SelectK,RandomK}Launcher) and by the fused set-conv / cost-volume
// kernels, so that both produce the same neighbour sets bit for bit.
//
// Semantics restated from the reference kernels (file:line relative to /root/reference):
//   window walk      tf_ops/2d_conv_select_k/fused_conv_g.cu:73-104, 2d_conv_random_k/fused_conv_g.cu:74-111
//   random-K accept  tf_ops/2d_conv_random_k/fused_conv_g.cu:126-150
//   select-K sort    tf_ops/2d_conv_select_k/fused_conv_g.cu:148-204
// The reference runs one *thread* per query over the window serially; here the 32 lanes of a warp
// test 32 window cells at a time and agree on slots with ballots / prefix popcounts, which gives
// the scan-order result of the serial loop without walking it serially.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elo {

constexpr unsigned FULL_MASK = 0xffffffffu;

// a*a + b*b + c*c exactly as nvcc contracts it in the reference build (SASS: FMUL, FFMA, FFMA).
// Spelled with rounding intrinsics so no compiler flag can change the sequence.
__device__ __forceinline__ float sq3(float x, float y, float z)
{
    return __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
}

struct Window {
    int h2, w2;              // extent of the searched grid (xyz2)
    int kH, kW, kt;          // window, kt = kH * kW
    int stride_h, stride_w;  // query (h, w) -> window centre (h / stride_h, w / stride_w)
    int K, flag_copy;
    float d2max;             // distance * distance
};

// Per-CTA table of window offsets in scan order: off[j] = (p / kW - kH/2, p % kW - kW/2), p = random_hw[j].
__device__ __forceinline__ void build_offsets(int2* off, const int* __restrict__ random_hw,
                                              int kt, int kH, int kW, int nthreads)
{
    const int hh = kH / 2, hw = kW / 2;
    for (int j = threadIdx.x; j < kt; j += nthreads) {
        const int p = __ldg(random_hw + j);
        off[j] = make_int2(p / kW - hh, p % kW - hw);
    }
}

struct Cand {
    bool valid;  // in-bounds, non-empty pixel            (counts into valid_idx)
    bool acc;    // ... and within `distance` of the centre (counts into valid_in_dis_idx)
    float d;     // max(|c - q|^2, 1e-10)
    int hh, ww;  // cell of the searched grid
};

__device__ __forceinline__ Cand eval_candidate(const float* __restrict__ g2, int2 o, int ch, int cw,
                                               const Window& g, float xc, float yc, float zc)
{
    Cand r;
    r.valid = false; r.acc = false; r.d = 0.f;
    int hh = ch + o.x, ww = cw + o.y;
    if (ww < 0) ww += g.w2;            // cylindrical wrap, once (reference :88-96)
    if (ww >= g.w2) ww -= g.w2;
    r.hh = hh; r.ww = ww;
    // rows are clipped (reference :82-86).  The reference would read out of bounds when kW/2 > w2
    // (a second wrap would be needed); such cells are skipped here instead.
    if (hh >= 0 && hh < g.h2 && ww >= 0 && ww < g.w2) {
        const float* q = g2 + ((size_t)hh * g.w2 + ww) * 3;
        const float xq = __ldg(q), yq = __ldg(q + 1), zq = __ldg(q + 2);
        if (!(sq3(xq, yq, zq) <= 1e-10f)) {           // FSETP.GTU in the reference: NaN counts as valid
            r.valid = true;
            const float d = fmaxf(sq3(__fsub_rn(xc, xq), __fsub_rn(yc, yq), __fsub_rn(zc, zq)), 1e-10f);
            r.d = d;
            r.acc = !(d > g.d2max);
        }
    }
    return r;
}

struct SearchCounts {
    int nvalid;  // leading ones of valid_idx
    int nsel;    // leading ones of valid_in_dis_idx (random-K: = number of slots filled, <= K)
    int first;   // random-K: packed (hh << 16 | ww) of the first accepted cell (for flag_copy), else 0
};

__device__ __forceinline__ int pack_hw(int hh, int ww) { return (hh << 16) | ww; }

// position of the n-th (1-based) set bit of m; m has at least n bits set
__device__ __forceinline__ int nth_set_bit(unsigned m, int n)
{
    for (int i = 1; i < n; ++i) m &= m - 1u;
    return __ffs(m) - 1;
}

// ---- random-K ------------------------------------------------------------------------------
// emit(slot, hh, ww) is called by the lane that owns the slot-th accepted cell, slot < K.
// Stops, like the reference's `break`, in the round where the K-th cell is accepted; valid cells
// after it in scan order are not counted.
template <typename Emit>
__device__ __forceinline__ SearchCounts search_random_k(const float* __restrict__ g2, const int2* off,
                                                        const Window& g, int ch, int cw, float xc,
                                                        float yc, float zc, Emit emit)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    SearchCounts c; c.nvalid = 0; c.nsel = 0; c.first = 0;
    for (int base = 0; base < g.kt; base += 32) {
        const int j = base + lane;
        Cand r; r.valid = false; r.acc = false; r.hh = 0; r.ww = 0;
        if (j < g.kt) r = eval_candidate(g2, off[j], ch, cw, g, xc, yc, zc);
        const unsigned bv = __ballot_sync(FULL_MASK, r.valid);
        const unsigned ba = __ballot_sync(FULL_MASK, r.acc);
        const int slot = c.nsel + __popc(ba & lt);
        if (r.acc && slot < g.K) emit(slot, r.hh, r.ww);
        if (c.nsel == 0 && ba != 0u) {
            const int src = __ffs(ba) - 1;
            c.first = __shfl_sync(FULL_MASK, pack_hw(r.hh, r.ww), src);
        }
        const int na = __popc(ba);
        if (c.nsel + na >= g.K) {
            // lane of the K-th accepted cell overall = (K - nsel)-th set bit of ba
            const int last = nth_set_bit(ba, g.K - c.nsel);
            const unsigned upto = (last >= 31) ? FULL_MASK : ((2u << last) - 1u);
            c.nvalid += __popc(bv & upto);
            c.nsel = g.K;
            return c;
        }
        c.nvalid += __popc(bv);
        c.nsel += na;
    }
    return c;
}

// ---- select-K ------------------------------------------------------------------------------
// dist[kt] / hw[kt] are per-warp scratch (shared memory).  The reference keeps these arrays in
// per-thread local memory and runs K steps of a selection sort with a *swap*, which makes the
// order of equal distances depend on the displaced elements.  To stay bit-exact
// in every tie case the same K steps are replayed here on the same array; only the arg-min scan
// of each step is spread over the lanes ((dist, position) lexicographic == "first minimum").
// emit(slot, hh, ww) is called from lane 0 for every written slot, in slot order.
// Returns counts (.first = entry 0 after the first step, i.e. what flag_copy duplicates);
// *nwritten = number of leading slots that were emitted.
template <typename Emit>
__device__ __forceinline__ SearchCounts search_select_k(const float* __restrict__ g2, const int2* off,
                                                        const Window& g, int ch, int cw, float xc,
                                                        float yc, float zc, float* dist, int* hw,
                                                        int* nwritten, Emit emit)
{
    const int lane = threadIdx.x & 31;
    SearchCounts c; c.nvalid = 0; c.nsel = 0; c.first = 0;
    for (int base = 0; base < g.kt; base += 32) {
        const int j = base + lane;
        Cand r; r.valid = false; r.acc = false; r.hh = 0; r.ww = 0; r.d = 0.f;
        if (j < g.kt) {
            r = eval_candidate(g2, off[j], ch, cw, g, xc, yc, zc);
            dist[j] = r.acc ? r.d : 1e10f;
            hw[j] = r.acc ? pack_hw(r.hh, r.ww) : 0;
        }
        c.nvalid += __popc(__ballot_sync(FULL_MASK, r.valid));
        c.nsel += __popc(__ballot_sync(FULL_MASK, r.acc));
    }
    __syncwarp();

    int written = 0;
    const int rounds = g.K < g.kt ? g.K : g.kt;
    for (int s = 0; s < rounds; ++s) {
        // first minimum of dist[s .. kt)
        unsigned bd = 0xffffffffu;   // +NaN pattern: larger than any finite key
        int bt = 0x7fffffff;
        for (int t = s + lane; t < g.kt; t += 32) {
            const float d = dist[t];
            // positive floats order like their bit patterns; keys here are >= 1e-10 or NaN-free
            const unsigned u = __float_as_uint(d);
            if (u < bd) { bd = u; bt = t; }
        }
        const unsigned md = __reduce_min_sync(FULL_MASK, bd);
        const int m = (int)__reduce_min_sync(FULL_MASK, bd == md ? (unsigned)bt : 0x7fffffffu);
        if (lane == 0) {
            const float dm = dist[m]; const int pm = hw[m];
            if (m != s) { dist[m] = dist[s]; hw[m] = hw[s]; dist[s] = dm; hw[s] = pm; }
            if (s == 0) c.first = pm;
            if (dm < 1e10f) emit(s, pm >> 16, pm & 0xffff);
        }
        __syncwarp();
        const float ds = __uint_as_float(md);
        if (s == 0) c.first = __shfl_sync(FULL_MASK, c.first, 0);
        if (!(ds < 1e10f)) break;    // everything left is a dummy (or >= 1e10): nothing more is written
        ++written;
    }
    *nwritten = written;
    return c;
}

// ============================================================================================
// Register-resident fast paths (kt <= 32 * S, S compile-time).  Same results as the functions above.
// All S window cells of a lane are evaluated first -- their global loads are independent and in flight
// together, one memory latency per query instead of one per 32 cells -- then the ballots / selection run
// on registers.

template <int S>
__device__ __forceinline__ void eval_all(const float* __restrict__ g2, const int2* off, const Window& g, int ch,
                                         int cw, float xc, float yc, float zc, bool (&val)[S], bool (&acc)[S],
                                         unsigned (&key)[S], int (&hw)[S])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const int j = lane + 32 * i;
        val[i] = false; acc[i] = false; key[i] = 0xffffffffu; hw[i] = 0;
        if (j < g.kt) {
            const Cand r = eval_candidate(g2, off[j], ch, cw, g, xc, yc, zc);
            val[i] = r.valid; acc[i] = r.acc;
            if (r.acc) { key[i] = __float_as_uint(r.d); hw[i] = pack_hw(r.hh, r.ww); }
        }
    }
}

template <int S, typename Emit>
__device__ __forceinline__ SearchCounts search_random_k_fast(const float* __restrict__ g2, const int2* off,
                                                             const Window& g, int ch, int cw, float xc, float yc,
                                                             float zc, Emit emit)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    bool val[S], acc[S];
    unsigned key[S];
    int hw[S];
    eval_all<S>(g2, off, g, ch, cw, xc, yc, zc, val, acc, key, hw);
    SearchCounts c; c.nvalid = 0; c.nsel = 0; c.first = 0;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        if (32 * i >= g.kt) break;
        const unsigned bv = __ballot_sync(FULL_MASK, val[i]);
        const unsigned ba = __ballot_sync(FULL_MASK, acc[i]);
        const int slot = c.nsel + __popc(ba & lt);
        if (acc[i] && slot < g.K) emit(slot, hw[i] >> 16, hw[i] & 0xffff);
        if (c.nsel == 0 && ba != 0u) c.first = __shfl_sync(FULL_MASK, hw[i], __ffs(ba) - 1);
        const int na = __popc(ba);
        if (c.nsel + na >= g.K) {
            const int last = nth_set_bit(ba, g.K - c.nsel);
            const unsigned upto = (last >= 31) ? FULL_MASK : ((2u << last) - 1u);
            c.nvalid += __popc(bv & upto);
            c.nsel = g.K;
            return c;
        }
        c.nvalid += __popc(bv);
        c.nsel += na;
    }
    return c;
}

// select-K on register keys.  Valid when K <= 32 and distance^2 < 1e10 (every accepted key is a real,
// finite distance).  K rounds of "smallest (distance, scan position)"; this equals the reference's
// swap-based selection sort whenever the K+1 smallest distances are pairwise different.  If two equal
// distances meet among them, *tie is set and NOTHING has been emitted: the caller replays the exact
// procedure (search_select_k) for this query.
template <int S, typename Emit>
__device__ __forceinline__ SearchCounts search_select_k_fast(const float* __restrict__ g2, const int2* off,
                                                             const Window& g, int ch, int cw, float xc, float yc,
                                                             float zc, int* nwritten, bool* tie, Emit emit)
{
    const int lane = threadIdx.x & 31;
    bool val[S], acc[S];
    unsigned key[S];
    int hw[S];
    eval_all<S>(g2, off, g, ch, cw, xc, yc, zc, val, acc, key, hw);
    SearchCounts c; c.nvalid = 0; c.nsel = 0; c.first = 0;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        c.nvalid += __popc(__ballot_sync(FULL_MASK, val[i]));
        c.nsel += __popc(__ballot_sync(FULL_MASK, acc[i]));
    }
    const int rounds = c.nsel < g.K ? c.nsel : g.K;
    unsigned prev = 0u;
    bool t = false;
    int res = 0;
    for (int s = 0; s < rounds; ++s) {
        unsigned bd = key[0];
        int bi = 0;
#pragma unroll
        for (int i = 1; i < S; ++i)
            if (key[i] < bd) { bd = key[i]; bi = i; }
        const unsigned md = __reduce_min_sync(FULL_MASK, bd);
        const unsigned js = __reduce_min_sync(FULL_MASK, bd == md ? (unsigned)(bi * 32 + lane) : 0x7fffffffu);
        t = t || (s > 0 && md == prev);
        prev = md;
        const int owner = js & 31, is = js >> 5;
        int mine = hw[0];
#pragma unroll
        for (int i = 1; i < S; ++i)
            if (i == is) mine = hw[i];
        const int sel = __shfl_sync(FULL_MASK, mine, owner);
        if (lane == s) res = sel;
        if (lane == owner) {
#pragma unroll
            for (int i = 0; i < S; ++i)
                if (i == is) key[i] = 0xffffffffu;
        }
    }
    if (c.nsel > rounds && rounds > 0) {          // is the first cell left out as close as the last one taken?
        unsigned bd = key[0];
#pragma unroll
        for (int i = 1; i < S; ++i) bd = min(bd, key[i]);
        t = t || (__reduce_min_sync(FULL_MASK, bd) == prev);
    }
    *tie = t;
    *nwritten = rounds;
    c.first = __shfl_sync(FULL_MASK, res, 0);    // entry 0 after the first step (all-dummy array: cell (0,0))
    if (!t && lane < rounds) emit(lane, res >> 16, res & 0xffff);
    return c;
}

// kt -> smallest supported S with 32 * S >= kt (0: use the generic path)
__host__ __device__ inline int fast_slots(int kt)
{
    const int s = (kt + 31) / 32;
    if (s <= 1) return 1;
    if (s <= 2) return 2;
    if (s <= 3) return 3;
    if (s <= 4) return 4;
    if (s <= 6) return 6;
    if (s <= 8) return 8;
    if (s <= 12) return 12;
    if (s <= 16) return 16;
    return 0;
}

// One query, fast path when possible, exact replay on ties.  Returns counts; emit semantics as above.
template <bool SELECT, typename Emit>
__device__ __forceinline__ SearchCounts search_query(const float* __restrict__ g2, const int2* off, const Window& g,
                                                     int ch, int cw, float xc, float yc, float zc, float* dist,
                                                     int* hw, int* nwritten, bool fast_ok, Emit emit)
{
    const int S = fast_ok ? fast_slots(g.kt) : 0;
    if (!SELECT) {
        SearchCounts c;
        switch (S) {
            case 1: c = search_random_k_fast<1>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 2: c = search_random_k_fast<2>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 3: c = search_random_k_fast<3>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 4: c = search_random_k_fast<4>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 6: c = search_random_k_fast<6>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 8: c = search_random_k_fast<8>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 12: c = search_random_k_fast<12>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            case 16: c = search_random_k_fast<16>(g2, off, g, ch, cw, xc, yc, zc, emit); break;
            default: c = search_random_k(g2, off, g, ch, cw, xc, yc, zc, emit); break;
        }
        *nwritten = c.nsel;
        return c;
    }
    bool tie = true;
    SearchCounts c;
    switch (S) {
        case 1: c = search_select_k_fast<1>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 2: c = search_select_k_fast<2>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 3: c = search_select_k_fast<3>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 4: c = search_select_k_fast<4>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 6: c = search_select_k_fast<6>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 8: c = search_select_k_fast<8>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 12: c = search_select_k_fast<12>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        case 16: c = search_select_k_fast<16>(g2, off, g, ch, cw, xc, yc, zc, nwritten, &tie, emit); break;
        default: break;
    }
    if (tie) {   // equal distances among the nearest (or no fast path): the reference's exact procedure
        c = search_select_k(g2, off, g, ch, cw, xc, yc, zc, dist, hw, nwritten, emit);
        __syncwarp();
    }
    return c;
}

}  // namespace elo
