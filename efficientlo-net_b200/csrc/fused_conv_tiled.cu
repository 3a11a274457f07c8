// fused_conv_tiled.cu -- tile-staged, thread-per-query form of the stand-alone neighbour-search ops
// (large query counts: BASELINE.json configs[0], every pixel of a 64x1800 frame as a query).
//
// Same results, bit for bit, as fused_conv_index.cu (and therefore as the reference kernels
// tf_ops/2d_conv_select_k/fused_conv_g.cu:11-209, tf_ops/2d_conv_random_k/fused_conv_g.cu:13-156); the
// difference is the work decomposition.  The warp-per-query kernel spends ~1600 warp instructions per
// query (K rounds of a warp-wide arg-min) and is issue-bound at 0.10 of the HBM roofline.  Here:
//   * a CTA owns TQ = 64 consecutive queries, one THREAD per query (like the reference) -- but
//   * the searched grid's neighbourhood of the CTA (bounding box of the 64 window centres + the window
//     halo, columns wrapped around the cylinder, rows outside the image as empty pixels) is staged ONCE in
//     shared memory as three planes, so the inner loop is 3 conflict-free LDS + 9 FP32 ops per window cell
//     with no wrap / bounds logic;
//   * select-K keeps the K+1 smallest candidates as a sorted register array of PACKED keys (distance bits
//     with the walk position in the low mantissa bits), maintained with a branch-free min/max chain.
//     Candidates that beat the current (K+1)-th key are first pushed into a per-thread shared-memory
//     queue; the chain runs once per queued item of the slowest lane, every 16 window cells, so a warp pays
//     for max-over-lanes insertions instead of one insertion per window cell.  The window is walked
//     centre-out (counting sort of dw^2 + 4 dh^2, per CTA) so the queue empties quickly: the result of a
//     select-K does not depend on the walk order unless distances tie;
//   * two selected keys that agree in the bits left for the distance (a real tie or a near-tie) send that
//     query to the exact swap-based replay of elo_search.cuh (warp-cooperative, inside the same CTA), which
//     reproduces the reference's unstable tie order;
//   * all four outputs of the CTA (contiguous in HBM: 64 rows of each tensor) are written with 16-byte
//     stores whose values are decoded from per-query counts in shared memory -- the 2 x kt floats of
//     valid_idx / valid_in_dis_idx per query are 83 % of the op's bytes.
// A CTA whose queries are not spatially compact (arbitrary idx_n2), or that spans two samples, reads the
// grid through the read-only cache instead of the staged tile; results are identical.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "../../include/elo_b200.h"
#include "elo_common.cuh"
#include "elo_search.cuh"

namespace elo {

constexpr int TQ = 64;       // queries = threads per CTA
constexpr int QG = 16;       // window cells between two drains of the candidate queue
constexpr unsigned KEY_NONE = 0x7f000000u;   // larger than any accepted key (d <= distance^2 < 1e10)

struct TiledParams {
    int B, H, W, N;
    Window g;
    const float* xyz1;
    const float* xyz2;
    const int* idx_n2;
    const int* random_hw;
    int* out_idx;
    float* out_valid;
    float* out_vdis;
    float* out_mask;
    long long total;
    int tile_cap;        // cells the staged tile may hold
    int jbits;           // low key bits that carry the walk position
    int nbins;           // bins of the centre-out counting sort
    int vec_ok;          // every output pointer is 16-byte aligned
    unsigned magic_kt;   // ceil(2^32 / kt), ceil(2^32 / K): exact quotients for the writer's ranges
    unsigned magic_k;
};

struct TileGeom {
    int staged;          // 1: tile holds the neighbourhood; 0: read the grid directly
    int row0, col0;      // grid cell of tile cell (0, 0) (col0 may lie outside [0, w2): wrapped on load)
    int th, tw;
    int hmin, rmin;
    int b;
};

__device__ __forceinline__ int wrap_once(int ww, int w2)
{
    if (ww < 0) ww += w2;
    if (ww >= w2) ww -= w2;
    return ww;
}

__device__ __forceinline__ unsigned udiv_magic(unsigned e, unsigned magic, unsigned d, unsigned& rem)
{
    unsigned q = __umulhi(e, magic);
    rem = e - q * d;
    if (rem >= d) { rem -= d; ++q; }      // magic = 2^32 - 1 stands in for d = 1
    return q;
}

// One insertion into the ascending array a[0..KR): afterwards a holds the KR smallest of (a, x).
// Branch-free and without a serial chain: a'[i] = min(a[i], max(a[i-1], x)).
template <int KR>
__device__ __forceinline__ void chain_insert(unsigned (&a)[KR], unsigned x)
{
    unsigned carry = x;
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const unsigned m = max(a[i], x);
        a[i] = min(a[i], carry);
        carry = m;
    }
}

template <bool SELECT, int KR>
__global__ void __launch_bounds__(TQ) fused_conv_tiled_kernel(const TiledParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();
    const Window g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kt = g.kt, K = g.K;

    // ---- shared-memory carve-up ------------------------------------------------------------------
    unsigned char* sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* r = sp; sp += (bytes + 15) & ~size_t(15); return r; };
    int2* off_scan = reinterpret_cast<int2*>(take((size_t)kt * 8));          // (dh, dw) in reference scan order
    int* walk_pk = reinterpret_cast<int*>(take((size_t)kt * 4));             // (dh << 16 | dw & 0xffff) in walk order
    int* walk_to = reinterpret_cast<int*>(take((size_t)kt * 4));             // tile offset of the walk's j-th cell
    int* sel = reinterpret_cast<int*>(take((size_t)K * TQ * 4));             // sel[s][tid]: packed (hh, ww) of slot s
    unsigned* queue = reinterpret_cast<unsigned*>(take((size_t)QG * TQ * 4));
    int* s_nvalid = reinterpret_cast<int*>(take(TQ * 4));
    int* s_nsel = reinterpret_cast<int*>(take(TQ * 4));
    int* s_nwr = reinterpret_cast<int*>(take(TQ * 4));
    int* s_first = reinterpret_cast<int*>(take(TQ * 4));
    int* s_bcopy = reinterpret_cast<int*>(take(TQ * 4));                     // b << 1 | copy
    float4* s_ctr = reinterpret_cast<float4*>(take(TQ * 16));                // centre xyz (+ ch / cw as int bits)
    int* s_chw = reinterpret_cast<int*>(take(TQ * 4));
    int* s_ties = reinterpret_cast<int*>(take(TQ * 4));
    int* s_misc = reinterpret_cast<int*>(take(64));                          // reductions, tie count, geometry
    float* fb_dist = nullptr;
    int* fb_hw = nullptr;
    int* bins = nullptr;
    if (SELECT) {
        fb_dist = reinterpret_cast<float*>(take((size_t)(TQ / 32) * kt * 4));
        fb_hw = reinterpret_cast<int*>(take((size_t)(TQ / 32) * kt * 4));
        bins = reinterpret_cast<int*>(take((size_t)(p.nbins + 1) * 4));
    }
    float* tile = reinterpret_cast<float*>(take((size_t)p.tile_cap * 12));
    float* tx = tile;
    float* ty = tile + p.tile_cap;
    float* tz = tile + 2 * p.tile_cap;

    // ---- this thread's query -----------------------------------------------------------------------
    const long long q0 = (long long)blockIdx.x * TQ;
    const long long q = q0 + tid;
    const bool active = q < p.total;
    int b = 0, h = 0, w = 0;
    float xc = 0.f, yc = 0.f, zc = 0.f;
    bool cvalid = false;
    if (active) {
        b = (int)(q / p.N);
        h = __ldg(p.idx_n2 + q * 2);
        w = __ldg(p.idx_n2 + q * 2 + 1);
        if (h >= 0 && h < p.H && w >= 0 && w < p.W) {       // reference: out-of-range = UB; here: empty row
            const float* c = p.xyz1 + ((size_t)b * p.H * p.W + (size_t)h * p.W + w) * 3;
            xc = __ldg(c); yc = __ldg(c + 1); zc = __ldg(c + 2);
            cvalid = !(fmaxf(sq3(xc, yc, zc), 1e-10f) <= 1e-10f);   // reference :61-69
        }
    }
    const int ch = h / g.stride_h, cw = w / g.stride_w;

    // ---- tables: scan order, and (select-K) the centre-out walk -----------------------------------------
    const int hh2 = g.kH / 2, hw2 = g.kW / 2;
    if (tid < 16) s_misc[tid] = tid == 1 ? min(max(cw, 0), g.w2 - 1) : 0;   // [0] tie count, [1] reference column (query 0)
    if (SELECT)
        for (int i = tid; i <= p.nbins; i += TQ) bins[i] = 0;
    for (int j = tid; j < kt; j += TQ) {
        const int pp = __ldg(p.random_hw + j);
        off_scan[j] = make_int2(pp / g.kW - hh2, pp % g.kW - hw2);
    }
    __syncthreads();

    // ---- geometry of the CTA's neighbourhood --------------------------------------------------------------
    // columns are measured relative to the first query's centre and folded onto (-w2/2, w2/2], so a run of
    // queries that crosses the end of an image row is still one compact box on the cylinder
    const int cwref = s_misc[1];
    int rel = cw - cwref;
    if (2 * rel > g.w2) rel -= g.w2;
    else if (2 * rel < -g.w2) rel += g.w2;
    {
        const bool colok = !cvalid || (cw >= 0 && cw < g.w2);
        const int big = 1 << 30;
        int hmin = __reduce_min_sync(FULL_MASK, cvalid ? ch : big);
        int hmax = __reduce_max_sync(FULL_MASK, cvalid ? ch : -big);
        int rmin = __reduce_min_sync(FULL_MASK, cvalid ? rel : big);
        int rmax = __reduce_max_sync(FULL_MASK, cvalid ? rel : -big);
        int bmin = __reduce_min_sync(FULL_MASK, active ? b : big);
        int bmax = __reduce_max_sync(FULL_MASK, active ? b : -big);
        const int allok = __all_sync(FULL_MASK, colok);
        int* red = reinterpret_cast<int*>(queue);       // free until the walk starts
        if (lane == 0) {
            int* r = red + warp * 8;
            r[0] = hmin; r[1] = hmax; r[2] = rmin; r[3] = rmax; r[4] = bmin; r[5] = bmax; r[6] = allok;
        }
        __syncthreads();
        if (tid == 0) {
            TileGeom tg;
            int ok = 1;
            hmin = big; hmax = -big; rmin = big; rmax = -big; bmin = big; bmax = -big;
            for (int wv = 0; wv < TQ / 32; ++wv) {
                const int* r = red + wv * 8;
                hmin = min(hmin, r[0]); hmax = max(hmax, r[1]);
                rmin = min(rmin, r[2]); rmax = max(rmax, r[3]);
                bmin = min(bmin, r[4]); bmax = max(bmax, r[5]);
                ok &= r[6];
            }
            tg.staged = 0; tg.row0 = 0; tg.col0 = 0; tg.th = 0; tg.tw = 1; tg.hmin = 0; tg.rmin = 0;
            tg.b = bmin == big ? 0 : bmin;
            if (hmin <= hmax) {
                const long long th = (long long)hmax - hmin + g.kH, tw = (long long)rmax - rmin + g.kW;
                // one wrap must be enough for every window cell (the reference wraps once, :88-96)
                const bool wrap_ok = hw2 <= g.w2;
                if (ok && bmin == bmax && wrap_ok && th * tw <= p.tile_cap) {
                    tg.staged = 1; tg.th = (int)th; tg.tw = (int)tw; tg.hmin = hmin; tg.rmin = rmin;
                    tg.row0 = hmin - hh2; tg.col0 = cwref + rmin - hw2;
                }
            } else {
                tg.staged = -1;                       // no valid centre in this CTA: nothing to search
            }
            *reinterpret_cast<TileGeom*>(s_misc + 4) = tg;
        }
        __syncthreads();
    }
    const TileGeom tg = *reinterpret_cast<const TileGeom*>(s_misc + 4);
    const bool staged = tg.staged == 1;
    const int tw = tg.tw;

    if (tg.staged >= 0) {
        // ---- stage the tile -------------------------------------------------------------------------
        if (staged) {
            const float* g2 = p.xyz2 + (size_t)tg.b * g.h2 * g.w2 * 3;
            const int cells = tg.th * tw;
            int c0 = tg.col0 % g.w2;
            if (c0 < 0) c0 += g.w2;
            for (int i = tid; i < cells; i += TQ) {
                const int r = i / tw, c = i - r * tw;
                const int gr = tg.row0 + r;
                const int gc = (c0 + c) % g.w2;
                float x = 0.f, y = 0.f, z = 0.f;
                if (gr >= 0 && gr < g.h2) {
                    const float* s = g2 + ((size_t)gr * g.w2 + gc) * 3;
                    x = __ldg(s); y = __ldg(s + 1); z = __ldg(s + 2);
                }
                tx[i] = x; ty[i] = y; tz[i] = z;
            }
        }
        // ---- walk tables ----------------------------------------------------------------------------------
        if (SELECT) {
            // centre-out: counting sort of the window cells by dw^2 + 4 dh^2 (scaled into nbins bins)
            const int kmax = hw2 * hw2 + 4 * hh2 * hh2;
            int shift = 0;
            while ((kmax >> shift) >= p.nbins) ++shift;
            for (int c = tid; c < kt; c += TQ) {
                const int dh = c / g.kW - hh2, dw = c % g.kW - hw2;
                atomicAdd(&bins[((dw * dw + 4 * dh * dh) >> shift) + 1], 1);
            }
            __syncthreads();
            if (warp == 0) {       // inclusive scan of bins[1..nbins] -> bins[i] = first position of bin i
                int run = 0;
                for (int base = 1; base <= p.nbins; base += 32) {
                    const int i = base + lane;
                    int v = i <= p.nbins ? bins[i] : 0;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(FULL_MASK, v, d);
                        if (lane >= d) v += t;
                    }
                    if (i <= p.nbins) bins[i] = run + v;
                    run += __shfl_sync(FULL_MASK, v, 31);
                }
            }
            __syncthreads();
            for (int c = tid; c < kt; c += TQ) {
                const int r = c / g.kW, cc = c % g.kW;
                const int dh = r - hh2, dw = cc - hw2;
                const int pos = atomicAdd(&bins[(dw * dw + 4 * dh * dh) >> shift], 1);
                walk_pk[pos] = (dh << 16) | (dw & 0xffff);
                walk_to[pos] = r * tw + cc;
            }
        } else {
            for (int j = tid; j < kt; j += TQ) {
                const int2 o = off_scan[j];
                walk_pk[j] = (o.x << 16) | (o.y & 0xffff);
                walk_to[j] = (o.x + hh2) * tw + (o.y + hw2);
            }
        }
        __syncthreads();

        // ---- the walk -----------------------------------------------------------------------------------
        const float* g2 = p.xyz2 + (size_t)b * g.h2 * g.w2 * 3;
        int base = 0;
        if (staged && cvalid) base = (ch - tg.hmin) * tw + (rel - tg.rmin);
        int nvalid = 0, nsel = 0;
        auto cell_of = [&](int j) {
            const int pk = walk_pk[j];
            return pack_hw(ch + (pk >> 16), wrap_once(cw + (int)(short)(pk & 0xffff), g.w2));
        };

        // ST = std::true_type: the window cells come from the staged tile; false_type: straight from the grid
        auto walk = [&](auto ST) {
            constexpr bool STG = decltype(ST)::value;
            // cell j of the walk (tile offset `to` when staged) -> (valid, accepted, distance)
            auto eval = [&](int j, int to, bool& valid, bool& acc, float& d) {
                float xq, yq, zq;
                bool inb = true;
                if constexpr (STG) {
                    const int t = base + to;
                    xq = tx[t]; yq = ty[t]; zq = tz[t];
                } else {
                    const int pk = walk_pk[j];
                    const int hh = ch + (pk >> 16);
                    const int ww = wrap_once(cw + (int)(short)(pk & 0xffff), g.w2);
                    inb = hh >= 0 && hh < g.h2 && ww >= 0 && ww < g.w2;
                    xq = yq = zq = 0.f;
                    if (inb && cvalid) {
                        const float* s = g2 + ((size_t)hh * g.w2 + ww) * 3;
                        xq = __ldg(s); yq = __ldg(s + 1); zq = __ldg(s + 2);
                    }
                }
                valid = cvalid && inb && !(sq3(xq, yq, zq) <= 1e-10f);   // FSETP.GTU in the reference: NaN is valid
                d = fmaxf(sq3(__fsub_rn(xc, xq), __fsub_rn(yc, yq), __fsub_rn(zc, zq)), 1e-10f);
                acc = valid && !(d > g.d2max);
            };

            if constexpr (SELECT) {
                unsigned a[KR];
#pragma unroll
                for (int i = 0; i < KR; ++i) a[i] = KEY_NONE;
                const unsigned jmask = (1u << p.jbits) - 1u;
                unsigned thr = KEY_NONE;
                int cnt = 0;
                auto visit = [&](int j, int to) {
                    bool valid, acc; float d;
                    eval(j, to, valid, acc, d);
                    nvalid += valid; nsel += acc;
                    const unsigned key = (__float_as_uint(d) & ~jmask) | (unsigned)j;
                    if (acc && key < thr) { queue[cnt * TQ + tid] = key; ++cnt; }
                };
                for (int jb = 0; jb < kt; jb += QG) {
                    if (jb + QG <= kt) {
#pragma unroll
                        for (int j4 = 0; j4 < QG; j4 += 4) {
                            const int4 to = *reinterpret_cast<const int4*>(walk_to + jb + j4);
                            visit(jb + j4, to.x); visit(jb + j4 + 1, to.y); visit(jb + j4 + 2, to.z); visit(jb + j4 + 3, to.w);
                        }
                    } else {
                        for (int j = jb; j < kt; ++j) visit(j, walk_to[j]);
                    }
                    // drain: one insertion per queued key of the slowest lane
                    const int n = __reduce_max_sync(FULL_MASK, cnt);
                    for (int t = 0; t < n; ++t) {
                        const unsigned x = t < cnt ? queue[t * TQ + tid] : 0xffffffffu;
                        chain_insert<KR>(a, x);
                    }
                    cnt = 0;
                    thr = a[KR - 1];
                }
                // near-ties among the K nearest (and against the first one left out) -> exact replay
                const int nw = min(nsel, K);
                bool tie = false;
#pragma unroll
                for (int i = 0; i + 1 < KR; ++i)
                    if (i < K && i + 1 < nsel && ((a[i] ^ a[i + 1]) & ~jmask) == 0u) tie = true;
                int first = 0;
                if (!tie) {
#pragma unroll
                    for (int s = 0; s < KR - 1; ++s)
                        if (s < nw) {
                            const int c = cell_of((int)(a[s] & jmask));
                            sel[s * TQ + tid] = c;
                            if (s == 0) first = c;
                        }
                } else {
                    s_ties[atomicAdd(&s_misc[0], 1)] = tid;
                }
                s_nwr[tid] = nw;
                s_first[tid] = first;
            } else {
                bool done = !cvalid;
                int first = 0;
                for (int j = 0; j < kt; ++j) {
                    if (!done) {
                        bool valid, acc; float d;
                        eval(j, walk_to[j], valid, acc, d);
                        nvalid += valid;
                        if (acc) {
                            const int c = cell_of(j);
                            if (nsel == 0) first = c;
                            sel[nsel * TQ + tid] = c;
                            ++nsel;
                            done = nsel >= K;                 // the reference's break (:149-150)
                        }
                    }
                    if (__all_sync(FULL_MASK, done)) break;
                }
                s_nwr[tid] = nsel;
                s_first[tid] = first;
            }
        };
        if (staged) walk(std::true_type{});
        else walk(std::false_type{});
        s_nvalid[tid] = nvalid;
        s_nsel[tid] = nsel;
        // select-K duplicates entry 0 even when nothing was in range (mask 1, index (b,0,0)); random-K only
        // once a first neighbour was accepted (reference select :180-192, random :126-138)
        s_bcopy[tid] = (b << 1) | ((cvalid && g.flag_copy == 1 && (SELECT || nsel > 0)) ? 1 : 0);
        if (SELECT) {
            s_ctr[tid] = make_float4(xc, yc, zc, 0.f);
            s_chw[tid] = pack_hw(ch, cw);
        }
        __syncthreads();

        // ---- exact replay of the tied queries (warp-cooperative, reference scan order) ---------------------
        if (SELECT) {
            const int nties = s_misc[0];
            float* dist = fb_dist + (size_t)warp * kt;
            int* hwv = fb_hw + (size_t)warp * kt;
            for (int t = warp; t < nties; t += TQ / 32) {
                const int qt = s_ties[t];
                const float4 c = s_ctr[qt];
                const int chw = s_chw[qt];
                const int bq = s_bcopy[qt] >> 1;
                const float* g2q = p.xyz2 + (size_t)bq * g.h2 * g.w2 * 3;
                auto emit = [&](int slot, int hh, int ww) { sel[slot * TQ + qt] = pack_hw(hh, ww); };
                int written = 0;
                const SearchCounts sc = search_select_k(g2q, off_scan, g, chw >> 16, chw & 0xffff, c.x, c.y, c.z, dist,
                                                        hwv, &written, emit);
                __syncwarp();
                if (lane == 0) { s_nwr[qt] = written; s_first[qt] = sc.first; }
                __syncwarp();
            }
            __syncthreads();
        }
    } else {
        s_nvalid[tid] = 0; s_nsel[tid] = 0; s_nwr[tid] = 0; s_first[tid] = 0; s_bcopy[tid] = b << 1;
        __syncthreads();
    }

    // ---- write the CTA's 64 rows of every output ------------------------------------------------------------
    const int nq = (int)min((long long)TQ, p.total - q0);
    // slot `sl` of the CTA (query sl / K, slot sl % K) -> (b, hh, ww) and mask
    auto slot_value = [&](unsigned sl, int& vb, int& vh, int& vw, float& vm) {
        unsigned s;
        const unsigned r = udiv_magic(sl, p.magic_k, (unsigned)K, s);
        const int bc = s_bcopy[r];
        int pk = 0;
        bool on = false;
        if ((int)s < s_nwr[r]) { pk = sel[s * TQ + r]; on = true; }
        else if (bc & 1) { pk = s_first[r]; on = true; }
        vb = on ? (bc >> 1) : 0; vh = pk >> 16; vw = pk & 0xffff; vm = on ? 1.0f : 0.0f;
    };
    {
        int* o_idx = p.out_idx + q0 * K * 3;
        float* o_mask = p.out_mask + q0 * K;
        const unsigned nslots = (unsigned)nq * K;
        const unsigned ngroups = p.vec_ok ? nslots / 4 : 0;
        for (unsigned gi = tid; gi < ngroups; gi += TQ) {
            int v[12];
            float m[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) slot_value(gi * 4 + i, v[3 * i], v[3 * i + 1], v[3 * i + 2], m[i]);
            int4* o = reinterpret_cast<int4*>(o_idx + (size_t)gi * 12);
            o[0] = make_int4(v[0], v[1], v[2], v[3]);
            o[1] = make_int4(v[4], v[5], v[6], v[7]);
            o[2] = make_int4(v[8], v[9], v[10], v[11]);
            reinterpret_cast<float4*>(o_mask)[gi] = make_float4(m[0], m[1], m[2], m[3]);
        }
        for (unsigned sl = ngroups * 4 + tid; sl < nslots; sl += TQ) {
            int vb, vh, vw; float vm;
            slot_value(sl, vb, vh, vw, vm);
            o_idx[(size_t)sl * 3] = vb; o_idx[(size_t)sl * 3 + 1] = vh; o_idx[(size_t)sl * 3 + 2] = vw;
            o_mask[sl] = vm;
        }
    }
    if (p.out_valid != nullptr || p.out_vdis != nullptr) {
        float* o_valid = p.out_valid ? p.out_valid + q0 * kt : nullptr;
        float* o_vdis = p.out_vdis ? p.out_vdis + q0 * kt : nullptr;
        const unsigned nel = (unsigned)nq * kt;
        const unsigned nvec = p.vec_ok ? nel / 4 : 0;
        for (unsigned v = tid; v < nvec; v += TQ) {
            unsigned pos;
            unsigned r = udiv_magic(v * 4, p.magic_kt, (unsigned)kt, pos);
            float a4[4], d4[4];
            int nv = s_nvalid[r], ns = s_nsel[r];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a4[i] = (int)pos < nv ? 1.0f : 0.0f;
                d4[i] = (int)pos < ns ? 1.0f : 0.0f;
                if (++pos == (unsigned)kt) {
                    pos = 0; ++r;
                    if (i < 3) { nv = s_nvalid[r & (TQ - 1)]; ns = s_nsel[r & (TQ - 1)]; }
                }
            }
            if (o_valid) reinterpret_cast<float4*>(o_valid)[v] = make_float4(a4[0], a4[1], a4[2], a4[3]);
            if (o_vdis) reinterpret_cast<float4*>(o_vdis)[v] = make_float4(d4[0], d4[1], d4[2], d4[3]);
        }
        for (unsigned e = nvec * 4 + tid; e < nel; e += TQ) {
            unsigned pos;
            const unsigned r = udiv_magic(e, p.magic_kt, (unsigned)kt, pos);
            if (o_valid) o_valid[e] = (int)pos < s_nvalid[r] ? 1.0f : 0.0f;
            if (o_vdis) o_vdis[e] = (int)pos < s_nsel[r] ? 1.0f : 0.0f;
        }
    }
}

static std::atomic<int> g_index_kernel{0};   // 0: by query count, 1: always tiled, 2: always warp-per-query

static unsigned magic_of(unsigned d)
{
    if (d <= 1) return 0xffffffffu;
    return (unsigned)(((1ull << 32) + d - 1) / d);
}

template <bool SELECT, int KR>
static cudaError_t launch_tiled_kr(const TiledParams& p, size_t smem, cudaStream_t stream)
{
    auto kern = fused_conv_tiled_kernel<SELECT, KR>;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    const long long ctas = (p.total + TQ - 1) / TQ;
    return launch(kern, dim3((unsigned)ctas), dim3(TQ), smem, stream, p);
}

// Returns 1 when the tiled kernel took the call (status in *rc), 0 when the caller should use the
// warp-per-query kernel (few queries, K > 32, distance^2 >= 1e10, window too large for shared memory).
int launch_index_tiled(bool select, int B, int H, int W, int N, const Window& g, const float* xyz1, const float* xyz2,
                       const int* idx_n2, const int* random_hw, int* out_idx, float* out_valid, float* out_vdis,
                       float* out_mask, cudaStream_t stream, int* rc)
{
    const long long total = (long long)B * N;
    const DeviceInfo& dev = device_info();
    const int force = g_index_kernel.load(std::memory_order_relaxed);
    if (force == 2) return 0;
    if (g.K > 32 || !(g.d2max < 1e10f)) return 0;
    if (force != 1 && total < (long long)dev.sm_count * TQ * 2) return 0;   // too few threads to fill the chip

    TiledParams p;
    p.B = B; p.H = H; p.W = W; p.N = N; p.g = g;
    p.xyz1 = xyz1; p.xyz2 = xyz2; p.idx_n2 = idx_n2; p.random_hw = random_hw;
    p.out_idx = out_idx; p.out_valid = out_valid; p.out_vdis = out_vdis; p.out_mask = out_mask;
    p.total = total;
    p.jbits = 1;
    while ((1 << p.jbits) < g.kt) ++p.jbits;
    p.nbins = 512;
    // raster-order queries: TQ centres on one or two rows (fewer columns when the window centre is strided)
    const long long cap = (long long)(g.kH + 1) * (TQ + g.kW);
    p.tile_cap = (int)cap;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vec_ok = aligned(out_idx) && aligned(out_valid) && aligned(out_vdis) && aligned(out_mask) ? 1 : 0;
    p.magic_kt = magic_of((unsigned)g.kt);
    p.magic_k = magic_of((unsigned)g.K);

    auto up = [](size_t b) { return (b + 15) & ~size_t(15); };
    size_t smem = up((size_t)g.kt * 8) + 2 * up((size_t)g.kt * 4) + up((size_t)g.K * TQ * 4) + up((size_t)QG * TQ * 4) +
                  5 * up(TQ * 4) + up(TQ * 16) + 2 * up(TQ * 4) + up(64);
    if (select) smem += 2 * up((size_t)(TQ / 32) * g.kt * 4) + up((size_t)(p.nbins + 1) * 4);
    smem += up((size_t)p.tile_cap * 12);
    if (smem > 96 * 1024) return 0;

    cudaError_t err;
    if (select) {
        if (g.K <= 6) err = launch_tiled_kr<true, 7>(p, smem, stream);
        else if (g.K <= 16) err = launch_tiled_kr<true, 17>(p, smem, stream);
        else err = launch_tiled_kr<true, 33>(p, smem, stream);
    } else {
        err = launch_tiled_kr<false, 1>(p, smem, stream);
    }
    *rc = err == cudaSuccess ? ELO_OK : set_cuda_error(err, select ? "fused_conv_select_k (tiled) launch"
                                                                  : "fused_conv_random_k (tiled) launch");
    return 1;
}

}  // namespace elo

extern "C" int elo_set_index_kernel(int which)
{
    if (which < 0 || which > 2) return elo::set_error(ELO_ERR_INVALID_ARGUMENT, "elo_set_index_kernel: 0, 1 or 2");
    elo::g_index_kernel.store(which, std::memory_order_relaxed);
    return ELO_OK;
}

extern "C" int elo_get_index_kernel(void) { return elo::g_index_kernel.load(std::memory_order_relaxed); }
