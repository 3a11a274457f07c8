// fused_conv_tiled.cu -- tile-staged, thread-per-query form of the stand-alone neighbour-search ops
// (large query counts: BASELINE.json configs[0], every pixel of a 64x1800 frame as a query).
//
// Same results, bit for bit, as fused_conv_index.cu (and therefore as the reference kernels
// tf_ops/2d_conv_select_k/fused_conv_g.cu:11-209, tf_ops/2d_conv_random_k/fused_conv_g.cu:13-156); the
// difference is the work decomposition.  The warp-per-query kernel spends ~1600 warp instructions per
// query (K rounds of a warp-wide arg-min) and is issue-bound at 0.10 of the HBM roofline.  Here:
//   * a CTA owns TQ = 128 / 160 / 192 consecutive queries, one THREAD per query (like the reference) -- but
//   * the searched grid's neighbourhood of the CTA (bounding box of the window centres + the window halo,
//     columns wrapped around the cylinder, rows outside the image as empty pixels) is staged ONCE in shared
//     memory as one float4 per cell (x, y, z, empty?), so a window cell costs one conflict-free LDS.128 and
//     ten FP32 operations, with no wrap / bounds / emptiness logic in the loop;
//   * select-K keeps the K+1 smallest candidates as a sorted register array of PACKED keys (distance bits
//     with the walk position in the low mantissa bits), maintained with a branch-free min/max chain.
//     Candidates that beat the current (K+1)-th key are first pushed into a per-thread shared-memory
//     queue; the chain runs once per queued item of the slowest lane, every 16 window cells, so a warp pays
//     for max-over-lanes insertions instead of one insertion per window cell.  The window is walked
//     centre-out (cells sorted by dw^2 + 4 dh^2 on the host, passed in the kernel parameters) so the queue
//     empties quickly: the result of a select-K does not depend on the walk order unless distances tie;
//   * two selected keys that agree in the bits left for the distance (a real tie or a near-tie) send that
//     query to the exact swap-based replay of elo_search.cuh (warp-cooperative, inside the same CTA), which
//     reproduces the reference's unstable tie order;
//   * all four outputs of the CTA (contiguous in HBM: TQ rows of each tensor) are written with 16-byte
//     stores whose values are decoded from per-query counts in shared memory -- the 2 x kt floats of
//     valid_idx / valid_in_dis_idx per query are 83 % of the op's bytes.
// A CTA whose queries are not spatially compact (arbitrary idx_n2), or that spans two samples, reads the
// grid through the read-only cache instead of the staged tile; results are identical.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <type_traits>
#include <vector>

#include "../../include/elo_b200.h"
#include "elo_bulk.cuh"
#include "elo_common.cuh"
#include "elo_count_rows.cuh"
#include "elo_search.cuh"

namespace elo {

// TQ queries = threads per CTA is a template parameter (128 / 160 / 192): the launcher picks the one that
// spreads the call's CTAs most evenly over the SMs in a single wave (115 200 queries: 720 CTAs of 160,
// five on 128 SMs and four on 20, instead of 900 CTAs of 128 with a seventh CTA on 12 SMs).
constexpr int QG = 16;       // window cells between two drains of the candidate queue
constexpr int LB = 8;        // staged cells loaded ahead of their use
constexpr int MAX_WALK = 1024;               // window cells whose host-sorted walk fits the kernel parameters
#ifdef ELO_TILED_TS      // experiment build only (tools/tiled_timeline.py): per-CTA phase timestamps
__device__ unsigned long long g_tiled_ts[1024 * 8];
__device__ __forceinline__ unsigned long long tiled_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TILED_TS(k, who) do { if (threadIdx.x == (who) && blockIdx.x < 1024) g_tiled_ts[blockIdx.x * 8 + (k)] = tiled_now(); } while (0)
#else
#define TILED_TS(k, who) do { } while (0)
#endif
constexpr unsigned KEY_NONE = 0x7f000000u;   // larger than any accepted key (d <= distance^2 < 1e10)
constexpr float EMPTY_X = 2e19f;             // x of an empty pixel in the staged tile: (x - EMPTY_X)^2 = +inf for |x| < 1e19

struct TiledParams {
    int B, H, W, N;
    Window g;
    const float* xyz1;
    const float* xyz2;
    const int* idx_n2;   // (B, N, 2) [h, w]; or NULL: the queries are the strided sub-grid `qs` (batched search)
    const int* random_hw;
    int* out_nbr;        // batched search: (B, N, K) linear cell of the searched grid or -1; then the four
                         // outputs of the op below are all NULL
    int qs_ow, qs_h, qs_w;   // sub-grid of the query image: query n -> cell ((n / ow) * qs_h, (n % ow) * qs_w)
    int* out_idx;
    float* out_valid;
    float* out_vdis;
    float* out_mask;
    long long total;
    int tile_cap;        // cells the staged tile may hold
    int tile_bytes;      // bytes of the tile region (also scratch of the exact replay)
    int jbits;           // low key bits that carry the walk position
    int vec_ok;          // every output pointer is 16-byte aligned
    unsigned magic_kt;   // ceil(2^32 / kt), ceil(2^32 / K): exact quotients for the writer's ranges
    unsigned magic_k;
    int dbg;             // timing aid (ELO_TILED_DBG): 1 = query warps skip the walk, 2 = store warp skips its rows, 4 = no replay
    int pitch;           // row pitch of the staged tile in cells (TQ + kW)
    float near_bound;    // sqrt(distance^2 / 12.5): coordinates within it are all mutually in range
    int bulk;            // 1: the tile's rows come in by bulk copies (xyz2 is 16-byte aligned); 0: plain loads
    int walk[MAX_WALK];  // select-K: window cells centre-out, (dh << 16) | (dw & 0xffff)
    int walk_to[MAX_WALK];  // ... and their byte offsets in the staged tile, (row * pitch + col) * 16
};

struct TileGeom {
    int staged;          // 1: tile holds the neighbourhood; 0: read the grid directly; -1: no valid centre
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

// One insertion into the ascending array a[0..KR): afterwards a holds the KR smallest of (a, x).
// Branch-free and without a serial chain: a'[i] = min(a[i], max(a[i-1], x)).  Returns the key that fell out.
template <int KR>
__device__ __forceinline__ unsigned chain_insert(unsigned (&a)[KR], unsigned x)
{
    unsigned carry = x;
#pragma unroll
    for (int i = 0; i < KR; ++i) {
        const unsigned m = max(a[i], x);
        a[i] = min(a[i], carry);
        carry = m;
    }
    return carry;
}

// Threads per CTA: TQ query threads; a select-K CTA has one more warp, the STORE WARP.  The counts of a select-K
// do not depend on the selection (the reference's walk visits all kt cells: valid_idx counts the non-empty in-image
// cells, valid_in_dis_idx those in range), and when every point of the neighbourhood is in range of every centre
// both are a box sum over the staged emptiness flags.  The store warp computes them right after the staging and
// streams out the CTA's 2 x TQ count rows (83 % of the op's bytes) WHILE the query warps walk their windows: the
// store-bound and the issue-bound halves of the op overlap, and the stores sit in the queue of a warp that has
// nothing else to do (stores issued early by the query threads themselves back-pressure their own shared-memory
// loads: DESIGN.md section 4.1).
template <bool SW, int TQ>
constexpr int tiled_threads() { return TQ + (SW ? 32 : 0); }
template <bool SELECT, int KR, int TQ, bool SW>
constexpr int tiled_min_ctas()
{
    // one wave of CTAs on 148 SMs for 115 200 queries needs 5 x 160 query threads per SM; with the store warp that
    // is 960 threads, i.e. 64 registers per thread instead of 72
    if (!SW) return KR <= 17 ? 896 / TQ : 512 / TQ;
    return KR <= 17 ? (TQ == 128 ? 6 : TQ == 160 ? 5 : 4) : 512 / tiled_threads<SW, TQ>();
}

template <bool SELECT, int KR, int TQ, bool SW>
__global__ void __launch_bounds__((tiled_threads<SW, TQ>()), (tiled_min_ctas<SELECT, KR, TQ, SW>()))
fused_conv_tiled_kernel(const __grid_constant__ TiledParams p)
{
    static_assert(SELECT || !SW, "only select-K has a store warp");
    constexpr int TW = TQ / 32;   // query warps per CTA
    constexpr int NT = tiled_threads<SW, TQ>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();
    TILED_TS(0, 0);
    const Window& g = p.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kt = g.kt, K = g.K;

    // ---- shared-memory carve-up ------------------------------------------------------------------
    unsigned char* sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* r = sp; sp += (bytes + 15) & ~size_t(15); return r; };
    int* walk_pk = reinterpret_cast<int*>(take((size_t)kt * 4));             // (dh << 16 | dw & 0xffff) in walk order
    int* walk_to = reinterpret_cast<int*>(take((size_t)kt * 4));             // byte offset of the walk's j-th cell in the tile
    // candidate queue during a select-K walk; afterwards sel[s][tid] = packed (hh, ww) of output slot s
    int* sel = reinterpret_cast<int*>(take((size_t)(K > QG ? K : QG) * TQ * 4));
    unsigned* queue = reinterpret_cast<unsigned*>(sel);
    float* s_nv = reinterpret_cast<float*>(take(TQ * 4));                    // leading ones of valid_idx
    float* s_ns = reinterpret_cast<float*>(take(TQ * 4));                    // ... of valid_in_dis_idx
    int* s_nwr = reinterpret_cast<int*>(take(TQ * 4));
    int* s_first = reinterpret_cast<int*>(take(TQ * 4));
    int* s_bcopy = reinterpret_cast<int*>(take(TQ * 4));                     // b << 1 | copy
    int* s_ties = reinterpret_cast<int*>(take(TQ * 4));
    int* s_qpos = reinterpret_cast<int*>(take(TQ * 4));                      // tile cell of the window's corner, or -1
    int* s_vsum = reinterpret_cast<int*>(take((size_t)(p.pitch + 1) * 4));   // store warp: prefix of the column sums
    int* s_misc = reinterpret_cast<int*>(take(64));                          // tie count, reference column, geometry
    unsigned char* tile = take((size_t)p.tile_bytes);                        // float4 per cell; replay scratch later
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 14);              // mbarrier of the tile's bulk copies

    // ---- this thread's query -----------------------------------------------------------------------
    const long long q0 = (long long)blockIdx.x * TQ;
    const long long q = q0 + tid;
    const bool store_warp = SW && tid >= TQ;
    const bool active = q < p.total && !store_warp;
    int b = 0, h = 0, w = 0;
    float xc = 0.f, yc = 0.f, zc = 0.f;
    bool cvalid = false;
    if (active) {
        b = (int)(q / p.N);
        if (p.idx_n2 != nullptr) {
            const int2 hw = __ldg(reinterpret_cast<const int2*>(p.idx_n2) + q);
            h = hw.x; w = hw.y;
        } else {
            const int n = (int)(q - (long long)b * p.N);
            h = (n / p.qs_ow) * p.qs_h; w = (n % p.qs_ow) * p.qs_w;
        }
        if (h >= 0 && h < p.H && w >= 0 && w < p.W) {       // reference: out-of-range = UB; here: empty row
            const float* c = p.xyz1 + ((size_t)b * p.H * p.W + (size_t)h * p.W + w) * 3;
            xc = __ldg(c); yc = __ldg(c + 1); zc = __ldg(c + 2);
            cvalid = !(fmaxf(sq3(xc, yc, zc), 1e-10f) <= 1e-10f);   // reference :61-69
        }
    }
    const int ch = h / g.stride_h, cw = w / g.stride_w;
    const int hh2 = g.kH / 2, hw2 = g.kW / 2;
    if (tid < 16) s_misc[tid] = tid == 1 ? min(max(cw, 0), g.w2 - 1) : 0;   // [0] tie count, [1] reference column
    __syncthreads();

    // ---- geometry of the CTA's neighbourhood --------------------------------------------------------------
    // columns are measured relative to one query's centre and folded onto [-w2/2, w2/2], so a run of
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
            for (int wv = 0; wv < TW; ++wv) {
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
                if (ok && bmin == bmax && wrap_ok && tw <= p.pitch && th * p.pitch <= p.tile_cap) {
                    tg.staged = 1; tg.th = (int)th; tg.tw = (int)tw; tg.hmin = hmin; tg.rmin = rmin;
                    tg.row0 = hmin - hh2; tg.col0 = cwref + rmin - hw2;
                }
            } else {
                tg.staged = -1;
            }
            *reinterpret_cast<TileGeom*>(s_misc + 4) = tg;
            mbar_init(s_bar, 32);            // the 32 lanes of warp 0 each announce the bytes of their tile rows
            mbar_fence_init();
        }
        __syncthreads();
    }
    TILED_TS(1, 0);
    const TileGeom tg = *reinterpret_cast<const TileGeom*>(s_misc + 4);
    const bool staged = tg.staged == 1;
    const int pitch = p.pitch;          // tile row pitch in cells: fixed per launch, so the walk's tile offsets
                                        // are launch constants (kernel parameters -> uniform registers)
    bool tied = false;                  // select-K: this query goes to the exact replay at the end

    // ---- row writers.  Each warp writes the 32 rows of its own queries: no CTA-wide barrier between walk and
    //      write, so the warps of an SM drift apart and one warp's stores overlap another's arithmetic. -----------
    const int r_lo = warp * 32;                                           // first CTA-local row of this warp
    const int nrows = store_warp ? 0 : (int)max(0ll, min(32ll, p.total - q0 - r_lo));
    // valid_idx / valid_in_dis_idx rows: elo_count_rows.cuh (skipped when the caller passed NULL for both, e.g. when
    // the stand-alone count kernel of fused_conv_counts.cu writes them concurrently)
    // (computed where they are used: kept live across the walk they would be spilled, and a reload from local memory
    // queues behind the stores in flight)
    auto o_valid = [&]() { return (p.out_valid && nrows > 0) ? p.out_valid + (q0 + r_lo) * kt : nullptr; };
    auto o_vdis = [&]() { return (p.out_vdis && nrows > 0) ? p.out_vdis + (q0 + r_lo) * kt : nullptr; };
    const bool count_rows = p.out_valid != nullptr || p.out_vdis != nullptr;      // CTA-uniform
    bool early_rows = false;            // CTA-uniform: the store warp writes the count rows while the others walk

    if (tg.staged >= 0) {
        // ---- stage the tile: (x, y, z, 1 if the pixel is empty) ---------------------------------------------
        // `far`: not every coordinate is within near_bound = sqrt(distance^2 / 12.5) of the origin.  When none
        // is, every point of the neighbourhood is within `distance` of every centre (|c - q|^2 <= 12 bound^2),
        // so valid_in_dis_idx == valid_idx and the walk needs neither the range test nor its counter.
        bool far = cvalid && !(fmaxf(fmaxf(fabsf(xc), fabsf(yc)), fabsf(zc)) <= p.near_bound);
        // ---- walk tables, and where each query's window starts in the tile (for the store warp) -----------------
        bool tables_done = false;
        auto fill_tables = [&]() {
            if (SELECT && !store_warp)
                s_qpos[tid] = (staged && cvalid) ? (((ch - tg.hmin) << 16) | (rel - tg.rmin)) : -1;
            for (int j = tid; j < kt; j += NT) {
                int pk, to;
                if (SELECT) {
                    pk = p.walk[j]; to = p.walk_to[j];
                } else {
                    const int pp = __ldg(p.random_hw + j);
                    const int r = pp / g.kW, cc = pp - r * g.kW;
                    pk = ((r - hh2) << 16) | ((cc - hw2) & 0xffff);
                    to = (r * pitch + cc) * 16;
                }
                walk_pk[j] = pk;
                walk_to[j] = to;
            }
            tables_done = true;
        };
        if (staged) {
            const float* g2 = p.xyz2 + (size_t)tg.b * g.h2 * g.w2 * 3;
            float4* t4 = reinterpret_cast<float4*>(tile);
            int c0 = tg.col0 % g.w2;
            if (c0 < 0) c0 += g.w2;
            constexpr int SR = 8;
            // empty pixel (reference :98-104); FSETP.GTU there: a NaN pixel counts as a point
            // An empty pixel is staged as (EMPTY_X, y, z, 1): its squared distance to any centre overflows, so the
            // select-K walk rejects it by its key alone
            auto put = [&](int r, int c, float x, float y, float z) {
                const bool empty = sq3(x, y, z) <= 1e-10f;
                t4[r * pitch + c] = make_float4(empty ? EMPTY_X : x, y, z, empty ? 1.0f : 0.0f);
                far = far || !(fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) <= p.near_bound);
            };
            if (p.bulk && tg.tw <= g.w2) {
                // ---- bulk-copy staging (TMA engine, UBLKCP): a tile row is one contiguous run of the grid, or two
                // when it wraps around the cylinder.  The runs land PACKED (12 bytes per point) at the head of the
                // row's own slot of the float4 tile -- start rounded down and length rounded up to four points, so
                // source, destination and size are multiples of 16 bytes -- and are then spread in place to
                // (x, y, z, empty) float4s, highest columns first so that no packed point is overwritten before it
                // has been read.  Rows outside the image take no copy; the last row of the tensor (whose rounded-up
                // run could end past the allocation) keeps plain loads.
                const int nA = min(tg.tw, g.w2 - c0);                 // cells before the wrap
                const long long rowcell0 = (long long)tg.b * g.h2 * g.w2;
                auto row_mode = [&](int gr) {                         // 0 outside the image, 1 bulk copy, 2 plain loads
                    return (gr < 0 || gr >= g.h2) ? 0 : (tg.b == p.B - 1 && gr == g.h2 - 1) ? 2 : 1;
                };
                // a run of n cells whose first cell has index = lead (mod 4) in the tensor: bytes of the copy
                auto run_bytes = [](int lead, int n) { return ((lead + n + 3) & ~3) * 12; };
                const int rc0m = (int)(rowcell0 & 3), w2m = g.w2 & 3;
                // la / lb: cells between the rounded-down start of the first / second run of grid row gr and its data
                auto leads = [&](int gr, int& la, int& lb) { lb = (rc0m + gr * w2m) & 3; la = (lb + c0) & 3; };
                if (warp == 0) {
                    unsigned bytes = 0;
                    for (int r = lane; r < tg.th; r += 32) {
                        const int gr = tg.row0 + r;
                        if (row_mode(gr) != 1) continue;
                        int la, lb;
                        leads(gr, la, lb);
                        bytes += run_bytes(la, nA) + (nA < tg.tw ? run_bytes(lb, tg.tw - nA) : 0);
                    }
                    if (bytes) mbar_expect_tx(s_bar, bytes); else mbar_arrive(s_bar);
                    for (int r = lane; r < tg.th; r += 32) {
                        const int gr = tg.row0 + r;
                        if (row_mode(gr) != 1) continue;
                        int la, lb;
                        leads(gr, la, lb);
                        const long long ra = rowcell0 + (long long)gr * g.w2;
                        unsigned char* slot = tile + (size_t)r * pitch * 16;
                        const int ba = run_bytes(la, nA);
                        bulk_g2s(slot, p.xyz2 + (ra + c0 - la) * 3, ba, s_bar);
                        if (nA < tg.tw) bulk_g2s(slot + ba, p.xyz2 + (ra - lb) * 3, run_bytes(lb, tg.tw - nA), s_bar);
                    }
                }
                fill_tables();               // under the copies' latency
                mbar_wait(s_bar, 0);
                // a packed point of column c sits below byte 12 (c + 9) of its slot and the float4 of column c' at
                // byte 16 c': columns >= c_lo may be written while columns < c_lo are still packed iff c_lo >= 27
                for (int r0 = 0; r0 < tg.th; r0 += SR) {
                    for (int c_hi = tg.tw; c_hi > 0;) {
                        int c_lo = c_hi - NT;
                        if (c_lo < 27) c_lo = c_hi <= NT ? 0 : 27;
                        const int c = c_lo + tid;
                        const bool mine = c < c_hi;
                        float x[SR], y[SR], z[SR];
#pragma unroll
                        for (int i = 0; i < SR; ++i) {
                            x[i] = y[i] = z[i] = 0.f;
                            const int r = r0 + i, gr = tg.row0 + r;
                            if (!mine || r >= tg.th) continue;
                            const int mode = row_mode(gr);
                            if (mode == 1) {
                                int la, lb;
                                leads(gr, la, lb);
                                const float* s = reinterpret_cast<const float*>(tile + (size_t)r * pitch * 16 +
                                                                                (c < nA ? 0 : run_bytes(la, nA))) +
                                                 (c < nA ? la + c : lb + c - nA) * 3;
                                x[i] = s[0]; y[i] = s[1]; z[i] = s[2];
                            } else if (mode == 2) {
                                const float* s = g2 + ((size_t)gr * g.w2 + (c0 + c) % g.w2) * 3;
                                x[i] = __ldg(s); y[i] = __ldg(s + 1); z[i] = __ldg(s + 2);
                            }
                        }
                        __syncthreads();
                        if (mine) {
#pragma unroll
                            for (int i = 0; i < SR; ++i)
                                if (r0 + i < tg.th) put(r0 + i, c, x[i], y[i], z[i]);
                        }
                        c_hi = c_lo;
                    }
                }
            } else {
                // a column of the tile per thread; its rows in groups of SR whose loads are all issued before any
                // of them is used (a plain row loop would pay one L2 round trip per row)
                for (int c = tid; c < tg.tw; c += NT) {
                    const int gc = (c0 + c) % g.w2;
                    for (int r0 = 0; r0 < tg.th; r0 += SR) {
                        float x[SR], y[SR], z[SR];
#pragma unroll
                        for (int i = 0; i < SR; ++i) {
                            const int gr = tg.row0 + r0 + i;
                            x[i] = y[i] = z[i] = 0.f;
                            if (r0 + i < tg.th && gr >= 0 && gr < g.h2) {
                                const float* s = g2 + ((size_t)gr * g.w2 + gc) * 3;
                                x[i] = __ldg(s); y[i] = __ldg(s + 1); z[i] = __ldg(s + 2);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < SR; ++i)
                            if (r0 + i < tg.th) put(r0 + i, c, x[i], y[i], z[i]);
                    }
                }
            }
        }
        if (far) s_misc[2] = 1;          // some coordinate of the neighbourhood is large, infinite or NaN
        if (!tables_done) fill_tables();
        __syncthreads();
        TILED_TS(2, 0);
        const bool all_near = SELECT && staged && s_misc[2] == 0;
        early_rows = SW && all_near && count_rows;
        if (store_warp) {
            if (early_rows) {
                // ---- store warp: counts of all TQ queries from the staged flags, then their 2 x TQ rows --------------
                const float4* t4 = reinterpret_cast<const float4*>(tile);
                for (int i = lane; i < TQ; i += 32) { s_nv[i] = 0.f; s_ns[i] = 0.f; }
                // one pass per window-row offset that occurs in the CTA (1 or 2 for raster-order queries): column sums
                // over the window's kH rows, exclusive prefix along the columns, then two look-ups per query
                for (int dr = 0; dr + g.kH <= tg.th; ++dr) {
                    int run = 0;
                    for (int cb = 0; cb < tg.tw; cb += 32) {
                        const int c = cb + lane;
                        int sc = 0;
                        if (c < tg.tw)
                            for (int r = 0; r < g.kH; ++r) sc += (int)t4[(dr + r) * pitch + c].w;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int nb = __shfl_up_sync(FULL_MASK, sc, o);
                            if (lane >= o) sc += nb;
                        }
                        if (c < tg.tw) s_vsum[c + 1] = run + sc;
                        run += __shfl_sync(FULL_MASK, sc, 31);
                    }
                    if (lane == 0) s_vsum[0] = 0;
                    __syncwarp();
                    for (int i = lane; i < TQ; i += 32) {
                        const int qp = s_qpos[i];
                        if (qp >= 0 && (qp >> 16) == dr) {
                            const int c0 = qp & 0xffff;
                            const float n = (float)(kt - (s_vsum[c0 + g.kW] - s_vsum[c0]));
                            s_nv[i] = n; s_ns[i] = n;
                        }
                    }
                    __syncwarp();
                }
                // all rows of the CTA in one go: the writer's per-column set-up is paid once per 32 float4 columns
                const int nr = (int)max(0ll, min((long long)TQ, p.total - q0));
                if (!(p.dbg & 2))
                    write_count_rows(p.out_valid ? p.out_valid + q0 * kt : nullptr, p.out_vdis ? p.out_vdis + q0 * kt : nullptr,
                                     nr, kt, p.vec_ok != 0, p.magic_kt, s_nv, s_ns, lane);
                TILED_TS(7, TQ);
            }
        } else {

        // ---- the walk -----------------------------------------------------------------------------------
        const float* g2 = p.xyz2 + (size_t)b * g.h2 * g.w2 * 3;
        const unsigned char* tb = tile;
        if (staged && cvalid) tb += ((ch - tg.hmin) * pitch + (rel - tg.rmin)) * 16;
        const float d2l = cvalid ? g.d2max : -1.0f;       // an invalid centre accepts nothing
        int nvalid = 0, nsel = 0;
        auto cell_of = [&](int j) {
            const int pk = walk_pk[j];
            return pack_hw(ch + (pk >> 16), wrap_once(cw + (int)(short)(pk & 0xffff), g.w2));
        };

        // ST = std::true_type: the window cells come from the staged tile; false_type: straight from the grid
        // NR = true_type (select-K on a staged tile only): every point is within range of every centre
        // CN = true_type: the counts of valid / in-range cells are wanted (count rows are written by this kernel)
        auto walk = [&](auto ST, auto NR, auto CN) {
            constexpr bool STG = decltype(ST)::value;
            constexpr bool NEAR = decltype(NR)::value;
            constexpr bool CNT = decltype(CN)::value;
            float ninv = 0.f;             // staged walk: empty cells, counted on the FP32 pipe
            int nrej = 0;                 // ... and cells that are empty or out of range
            // direct path: cell j of the walk -> accepted?, distance; counts the valid cells
            auto eval_direct = [&](int j, bool& acc, float& d) {
                const int pk = walk_pk[j];
                const int hh = ch + (pk >> 16);
                const int ww = wrap_once(cw + (int)(short)(pk & 0xffff), g.w2);
                const bool inb = hh >= 0 && hh < g.h2 && ww >= 0 && ww < g.w2;
                float xq = 0.f, yq = 0.f, zq = 0.f;
                if (inb && cvalid) {
                    const float* s = g2 + ((size_t)hh * g.w2 + ww) * 3;
                    xq = __ldg(s); yq = __ldg(s + 1); zq = __ldg(s + 2);
                }
                const bool valid = cvalid && inb && !(sq3(xq, yq, zq) <= 1e-10f);
                nvalid += valid;
                d = fmaxf(sq3(__fsub_rn(xc, xq), __fsub_rn(yc, yq), __fsub_rn(zc, zq)), 1e-10f);
                acc = valid && !(d > g.d2max);
            };
            // staged path: the cell is already in registers
            auto eval_cell = [&](const float4& c, bool& acc, float& d) {
                ninv += c.w;
                d = fmaxf(sq3(__fsub_rn(xc, c.x), __fsub_rn(yc, c.y), __fsub_rn(zc, c.z)), 1e-10f);
                acc = !(__fmaf_rn(c.w, 1e30f, d) > d2l);      // + 0 for a point, + 1e30 for an empty pixel
            };
            // exact distance of walk cell j again (near-tie fix-up)
            auto exact_d = [&](int j) {
                bool acc; float d;
                if constexpr (STG) {
                    const float4 c = *reinterpret_cast<const float4*>(tb + walk_to[j]);
                    d = fmaxf(sq3(__fsub_rn(xc, c.x), __fsub_rn(yc, c.y), __fsub_rn(zc, c.z)), 1e-10f);
                } else {
                    const int keep = nvalid;
                    eval_direct(j, acc, d);
                    nvalid = keep;
                }
                return d;
            };

            if constexpr (SELECT) {
                unsigned a[KR];
#pragma unroll
                for (int i = 0; i < KR; ++i) a[i] = KEY_NONE;
                const unsigned jmask = (1u << p.jbits) - 1u, nmask = ~jmask;     // jbits >= 4: QG positions fit
                unsigned thr = NEAR ? (KEY_NONE | jmask) : 0xffffffffu;
                unsigned ev = 0xffffffffu;      // smallest key that ever fell out of the array (a cell the filter turned
                                                // away had larger distance bits than the last key of its time)
                const unsigned qbase = (unsigned)__cvta_generic_to_shared(queue + tid);
                unsigned qp = qbase;
                // centre of a query that must accept nothing (NEAR path): every distance overflows
                const float xs = (NEAR && !cvalid) ? 3e38f : xc, ys = (NEAR && !cvalid) ? 3e38f : yc,
                            zs = (NEAR && !cvalid) ? 3e38f : zc;
                // The key pushed by the filter carries only the position inside the group (a compile-time
                // constant in the unrolled loop); the group base is OR-ed in when the queue is drained.  The
                // filter lets every key through whose distance bits are <= those of the current last key, so
                // the array always holds the KR smallest (distance bits, position) keys exactly.
                auto push = [&](int jj, bool acc, float d) {
                    const unsigned key = (__float_as_uint(d) & nmask) | (unsigned)jj;
                    if (acc && key <= thr) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qp), "r"(key) : "memory");
                        qp += TQ * 4;
                    }
                    if constexpr (!NEAR) {
                        if constexpr (STG) { if (!acc) ++nrej; } else { nsel += acc; }
                    }
                };
                // NEAR: an empty pixel (staged with x = EMPTY_X) gets a key above every threshold, so the key test is the
                // whole filter.  The reference's max(d, 1e-10) is left out of the loop: it only matters when two of a
                // query's candidates are that close, which is caught after the walk and sent to the exact replay.
                auto eval_near = [&](const float4& c, float& d) {
                    if constexpr (CNT) ninv += c.w;
                    d = sq3(__fsub_rn(xs, c.x), __fsub_rn(ys, c.y), __fsub_rn(zs, c.z));
                    return d;
                };
                for (int jb = 0; jb < kt; jb += QG) {
                    if (jb + QG <= kt) {
                        // the loads of LB cells are issued together, ahead of the queue stores the compiler
                        // would otherwise have to keep them behind (shared memory may alias)
#pragma unroll
                        for (int j8 = 0; j8 < QG; j8 += LB) {
                            if constexpr (STG) {
                                float4 c[LB];
#pragma unroll
                                for (int i = 0; i < LB; ++i) c[i] = *reinterpret_cast<const float4*>(tb + p.walk_to[jb + j8 + i]);
#pragma unroll
                                for (int i = 0; i < LB; ++i) {
                                    bool acc = true; float d;
                                    if constexpr (NEAR) { float dd; d = eval_near(c[i], dd); }
                                    else eval_cell(c[i], acc, d);
                                    push(j8 + i, acc, d);
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < LB; ++i) {
                                    bool acc; float d;
                                    eval_direct(jb + j8 + i, acc, d);
                                    push(j8 + i, acc, d);
                                }
                            }
                        }
                    } else {
                        for (int j = jb; j < kt; ++j) {
                            bool acc = true; float d;
                            if constexpr (NEAR) { float dd; d = eval_near(*reinterpret_cast<const float4*>(tb + walk_to[j]), dd); }
                            else if constexpr (STG) eval_cell(*reinterpret_cast<const float4*>(tb + walk_to[j]), acc, d);
                            else eval_direct(j, acc, d);
                            push(j - jb, acc, d);
                        }
                    }
                    // drain: one insertion per queued key of the slowest lane
                    const int cnt = (int)(qp - qbase) / (TQ * 4);
                    const int n = __reduce_max_sync(FULL_MASK, cnt);
                    for (int t = 0; t < n; ++t) {
                        const unsigned x = t < cnt ? (queue[t * TQ + tid] | (unsigned)jb) : 0xffffffffu;
                        ev = min(ev, chain_insert<KR>(a, x));
                    }
                    qp = qbase;
                    thr = a[KR - 1] | jmask;
                }
                if constexpr (NEAR && CNT) { nvalid = cvalid ? kt - (int)ninv : 0; nsel = nvalid; }
                if constexpr (NEAR && !CNT) {
                    // no count rows to write: all that is needed is how many keys are real, min(in-range cells, KR)
                    int nk = 0;
#pragma unroll
                    for (int i = 0; i < KR; ++i) nk += a[i] < KEY_NONE ? 1 : 0;
                    nsel = nk; nvalid = nk;
                }
                if constexpr (STG && !NEAR) { nsel = kt - nrej; nvalid = kt - (int)ninv; }
                if (!cvalid) { nsel = 0; nvalid = 0; }
                // Near-ties: adjacent keys whose distance bits agree.  An isolated pair -- inside the K nearest or
                // across the K-th place -- is put in order by its exact distances; an exact tie or a run of three
                // (counting the smallest key that fell out of the array) goes to the exact replay.
                const int nw = min(nsel, K);
                if constexpr (NEAR) {
                    // two candidates within the clamp of the reference's max(d, 1e-10): their order there is by scan
                    // position, not by the unclamped keys of this walk
                    if (nsel >= 2 && (a[1] & nmask) <= (__float_as_uint(1e-10f) & nmask)) tied = true;
                }
                bool prev_eq = false;
                // cheap screen first: the fix-up below is long straight-line code that most warps never need (two
                // of a query's K nearest agree in the distance bits of the key in ~1 % of the queries)
                bool any_eq = false;
#pragma unroll
                for (int i = 0; i + 1 < KR; ++i)
                    any_eq = any_eq || (i < K && i + 1 < nsel && ((a[i] ^ a[i + 1]) & nmask) == 0u);
                if (any_eq)
#pragma unroll
                for (int i = 0; i + 1 < KR; ++i) {
                    const bool eq = i < K && i + 1 < nsel && ((a[i] ^ a[i + 1]) & nmask) == 0u;
                    if (eq) {
                        // a pair that straddles the K-th place can be settled by its exact distances unless the next
                        // key -- in the array, or the smallest that fell out of it -- agrees with it in the distance bits
                        // too (that cell could belong between them)
                        const unsigned nxt = (i + 2 < KR) ? a[i + 2 < KR ? i + 2 : KR - 1] : ev;
                        if (prev_eq || (i + 1 >= K && ((nxt ^ a[i + 1]) & nmask) == 0u)) {
                            tied = true;
                        } else {
                            const float d0 = exact_d((int)(a[i] & jmask)), d1 = exact_d((int)(a[i + 1] & jmask));
                            if (d0 == d1) tied = true;
                            else if (d1 < d0) { const unsigned t = a[i]; a[i] = a[i + 1]; a[i + 1] = t; }
                        }
                    }
                    prev_eq = eq;
                }
                int first = 0;
#pragma unroll
                for (int s = 0; s < KR - 1; ++s)
                    if (s < nw) {
                        const int c = cell_of((int)(a[s] & jmask));
                        sel[s * TQ + tid] = c;
                        if (s == 0) first = c;
                    }
                if (tied) s_ties[atomicAdd(&s_misc[0], 1)] = tid;
                s_nwr[tid] = nw;
                s_first[tid] = first;
            } else {
                bool done = !cvalid;
                int first = 0;
                for (int j = 0; j < kt; ++j) {
                    if (!done) {
                        bool acc; float d;
                        if constexpr (STG) eval_cell(*reinterpret_cast<const float4*>(tb + walk_to[j]), acc, d);
                        else eval_direct(j, acc, d);
                        if (acc) {
                            const int c = cell_of(j);
                            if (nsel == 0) first = c;
                            sel[nsel * TQ + tid] = c;
                            ++nsel;
                            if (nsel >= K) {                  // the reference's break (:149-150): cells after
                                done = true;                  // this one are not counted
                                if constexpr (STG) nvalid = j + 1 - (int)ninv;
                            }
                        }
                    }
                    if (__all_sync(FULL_MASK, done)) break;
                }
                if constexpr (STG) {
                    if (!done) nvalid = kt - (int)ninv;
                }
                if (!cvalid) { nsel = 0; nvalid = 0; }
                s_nwr[tid] = nsel;
                s_first[tid] = first;
            }
        };
        // all_near with a store warp: the counts are not the walk's business -- the store warp has them
        if (p.dbg & 1) { s_nwr[tid] = 0; s_first[tid] = 0; }
        else if (all_near) {
            if (SW || !count_rows) walk(std::true_type{}, std::true_type{}, std::false_type{});
            else walk(std::true_type{}, std::true_type{}, std::true_type{});
        }
        else if (staged) walk(std::true_type{}, std::false_type{}, std::true_type{});
        else walk(std::false_type{}, std::false_type{}, std::true_type{});
        if (!early_rows) { s_nv[tid] = (float)nvalid; s_ns[tid] = (float)nsel; }
        // select-K duplicates entry 0 even when nothing was in range (mask 1, index (b,0,0)); random-K only
        // once a first neighbour was accepted (reference select :180-192, random :126-138)
        s_bcopy[tid] = (b << 1) | ((cvalid && g.flag_copy == 1 && (SELECT || nsel > 0)) ? 1 : 0);
        }   // query warps
    } else if (!store_warp) {
        s_nv[tid] = 0.f; s_ns[tid] = 0.f; s_nwr[tid] = 0; s_first[tid] = 0; s_bcopy[tid] = b << 1;
    }
    __syncwarp();
    TILED_TS(3, 0);

    // slot `sl` of the warp (row sl / K, slot sl % K) -> (b, hh, ww) and mask
    auto slot_value = [&](unsigned sl, int& vb, int& vh, int& vw, float& vm) {
        unsigned s;
        const unsigned r = r_lo + udiv_magic(sl, p.magic_k, (unsigned)K, s);
        const int bc = s_bcopy[r];
        int pk = 0;
        bool on = false;
        if ((int)s < s_nwr[r]) { pk = sel[s * TQ + r]; on = true; }
        else if (bc & 1) { pk = s_first[r]; on = true; }
        vb = on ? (bc >> 1) : 0; vh = pk >> 16; vw = pk & 0xffff; vm = on ? 1.0f : 0.0f;
    };
    if (nrows > 0 && p.out_nbr != nullptr) {
        // batched-search form: one int per slot, the warp's 32 rows are contiguous
        int* o_nbr = p.out_nbr + (q0 + r_lo) * K;
        const unsigned nslots = (unsigned)nrows * K;
        for (unsigned sl = lane; sl < nslots; sl += 32) {
            int vb, vh, vw; float vm;
            slot_value(sl, vb, vh, vw, vm);
            o_nbr[sl] = vm != 0.f ? vh * g.w2 + vw : -1;
        }
    }
    if (nrows > 0 && p.out_idx != nullptr) {
        int* o_idx = p.out_idx + (q0 + r_lo) * K * 3;
        float* o_mask = p.out_mask + (q0 + r_lo) * K;
        const unsigned nslots = (unsigned)nrows * K;
        const unsigned ngroups = p.vec_ok ? nslots / 4 : 0;
        for (unsigned gi = lane; gi < ngroups; gi += 32) {
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
        for (unsigned sl = ngroups * 4 + lane; sl < nslots; sl += 32) {
            int vb, vh, vw; float vm;
            slot_value(sl, vb, vh, vw, vm);
            o_idx[(size_t)sl * 3] = vb; o_idx[(size_t)sl * 3 + 1] = vh; o_idx[(size_t)sl * 3 + 2] = vw;
            o_mask[sl] = vm;
        }
    }
    TILED_TS(4, 0);
    if (!early_rows && !store_warp)
        write_count_rows(o_valid(), o_vdis(), nrows, kt, p.vec_ok != 0, p.magic_kt, s_nv + r_lo, s_ns + r_lo, lane);
    TILED_TS(5, 0);

    // ---- exact replay of the tied queries (warp-cooperative, reference scan order), then their rows again ---------
    if (SELECT) {
        __syncthreads();                        // every walk is over: tie list complete, tile free
        const int nties = (p.dbg & 4) ? 0 : s_misc[0];       // dbg 4: no replay (timing aid; tied queries stay approximate)
        if (nties > 0) {
            int2* off_scan = reinterpret_cast<int2*>(tile);
            float* dist = reinterpret_cast<float*>(tile + (size_t)kt * 8) + (size_t)warp * kt;
            int* hwv = reinterpret_cast<int*>(tile + (size_t)kt * 8 + (size_t)TW * kt * 4) + (size_t)warp * kt;
            build_offsets(off_scan, p.random_hw, kt, g.kH, g.kW, NT);
            __syncthreads();
            for (int t = warp; t < nties && !store_warp; t += TW) {
                const int qt = s_ties[t];
                const long long gq = q0 + qt;
                const int bq = (int)(gq / p.N);
                int2 hwq;
                if (p.idx_n2 != nullptr) {
                    hwq = __ldg(reinterpret_cast<const int2*>(p.idx_n2) + gq);
                } else {
                    const int n = (int)(gq - (long long)bq * p.N);
                    hwq = make_int2((n / p.qs_ow) * p.qs_h, (n % p.qs_ow) * p.qs_w);
                }
                const float* c = p.xyz1 + ((size_t)bq * p.H * p.W + (size_t)hwq.x * p.W + hwq.y) * 3;
                const float* g2q = p.xyz2 + (size_t)bq * g.h2 * g.w2 * 3;
                int* o_idx = p.out_idx ? p.out_idx + gq * K * 3 : nullptr;
                float* o_mask = p.out_mask ? p.out_mask + gq * K : nullptr;
                int* o_nbr = p.out_nbr ? p.out_nbr + gq * K : nullptr;
                auto emit = [&](int slot, int hh, int ww) {
                    if (o_nbr != nullptr) { o_nbr[slot] = hh * g.w2 + ww; return; }
                    o_idx[slot * 3] = bq; o_idx[slot * 3 + 1] = hh; o_idx[slot * 3 + 2] = ww;
                    o_mask[slot] = 1.0f;
                };
                int written = 0;
                const SearchCounts sc = search_select_k(g2q, off_scan, g, hwq.x / g.stride_h, hwq.y / g.stride_w, __ldg(c),
                                                        __ldg(c + 1), __ldg(c + 2), dist, hwv, &written, emit);
                __syncwarp();
                // slots the replay did not emit: zero, or the duplicate of entry 0 (flag_copy)
                const bool copy = g.flag_copy == 1;
                for (int k = written + lane; k < K; k += 32) {
                    if (o_nbr != nullptr) { o_nbr[k] = -1; continue; }
                    o_idx[k * 3] = copy ? bq : 0;
                    o_idx[k * 3 + 1] = copy ? sc.first >> 16 : 0;
                    o_idx[k * 3 + 2] = copy ? sc.first & 0xffff : 0;
                    o_mask[k] = copy ? 1.0f : 0.0f;
                }
                __syncwarp();
            }
        }
    }
    TILED_TS(6, 0);
}

#ifdef ELO_TILED_TS
extern "C" int elo_debug_tiled_ts(unsigned long long* host, int n)
{
    return (int)cudaMemcpyFromSymbol(host, g_tiled_ts, sizeof(unsigned long long) * (size_t)n);
}
#endif

static std::atomic<int> g_index_kernel{0};   // 0: by query count, 1: always tiled, 2: always warp-per-query
static std::atomic<int> g_store_warp_min_kt{getenv("ELO_STORE_WARP_KT") ? atoi(getenv("ELO_STORE_WARP_KT")) : 128};
// how the tile comes into shared memory: 0 = bulk copies (TMA engine), 1 = plain loads (elo_set_tile_staging)
static std::atomic<int> g_tile_staging{getenv("ELO_TILE_STAGING") ? atoi(getenv("ELO_TILE_STAGING")) : 0};

static unsigned magic_of(unsigned d)
{
    if (d <= 1) return 0xffffffffu;
    return (unsigned)(((1ull << 32) + d - 1) / d);
}

template <bool SELECT, int KR, int TQ, bool SW>
static cudaError_t launch_tiled_kr(const TiledParams& p, size_t smem, cudaStream_t stream)
{
    auto kern = fused_conv_tiled_kernel<SELECT, KR, TQ, SW>;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    const long long ctas = (p.total + TQ - 1) / TQ;
    return launch(kern, dim3((unsigned)ctas), dim3(tiled_threads<SW, TQ>()), smem, stream, p);
}

template <int TQ>
static cudaError_t launch_tiled_tq(bool select, bool sw, const TiledParams& p, size_t smem, cudaStream_t stream)
{
    if (!select) return launch_tiled_kr<false, 1, TQ, false>(p, smem, stream);
    if (sw) {
        if (p.g.K <= 6) return launch_tiled_kr<true, 7, TQ, true>(p, smem, stream);
        if (p.g.K <= 16) return launch_tiled_kr<true, 17, TQ, true>(p, smem, stream);
        return launch_tiled_kr<true, 33, TQ, true>(p, smem, stream);
    }
    if (p.g.K <= 6) return launch_tiled_kr<true, 7, TQ, false>(p, smem, stream);
    if (p.g.K <= 16) return launch_tiled_kr<true, 17, TQ, false>(p, smem, stream);
    return launch_tiled_kr<true, 33, TQ, false>(p, smem, stream);
}

static size_t tiled_smem(const Window& g, bool select, int tq, int* tile_cap, int* tile_bytes)
{
    auto up = [](size_t b) { return (b + 15) & ~size_t(15); };
    // raster-order queries: tq centres on one or two rows (fewer columns when the window centre is strided)
    *tile_cap = (g.kH + 1) * (tq + g.kW);      // rows x pitch
    size_t tb = (size_t)*tile_cap * 16;
    if (select) tb = std::max(tb, (size_t)g.kt * 8 + 2 * (size_t)(tq / 32) * g.kt * 4);   // replay scratch
    *tile_bytes = (int)up(tb);
    return 2 * up((size_t)g.kt * 4) + up((size_t)std::max(g.K, QG) * tq * 4) + 7 * up((size_t)tq * 4) +
           up((size_t)(tq + g.kW + 1) * 4) + up(64) + (size_t)*tile_bytes;
}

// Everything of TiledParams that follows from the window: key layout, writer constants, the centre-out walk.
static void tiled_prepare(TiledParams& p, bool select)
{
    const Window& g = p.g;
    p.jbits = 4;
    while ((1 << p.jbits) < g.kt) ++p.jbits;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.vec_ok = aligned(p.out_idx) && aligned(p.out_valid) && aligned(p.out_vdis) && aligned(p.out_mask) ? 1 : 0;
    p.magic_kt = magic_of((unsigned)g.kt);
    p.magic_k = magic_of((unsigned)g.K);
    p.dbg = getenv("ELO_TILED_DBG") ? atoi(getenv("ELO_TILED_DBG")) : 0;
    p.bulk = (g_tile_staging.load(std::memory_order_relaxed) == 0 && p.xyz2 != nullptr &&
              (reinterpret_cast<uintptr_t>(p.xyz2) & 15) == 0) ? 1 : 0;
    if (select) {
        // centre-out walk: nearest pixels first, rows weighted 2x (a LiDAR's rows are ~2x further apart than
        // its columns), so the K-th key tightens early and few later cells pass the filter
        const int hh2 = g.kH / 2, hw2 = g.kW / 2;
        std::vector<std::pair<int, int>> order((size_t)g.kt);
        for (int c = 0; c < g.kt; ++c) {
            const int dh = c / g.kW - hh2, dw = c % g.kW - hw2;
            order[c] = std::make_pair(dw * dw + 4 * dh * dh, c);
        }
        std::sort(order.begin(), order.end());
        for (int j = 0; j < g.kt; ++j) {
            const int c = order[j].second;
            p.walk[j] = ((c / g.kW - hh2) << 16) | ((c % g.kW - hw2) & 0xffff);
        }
    }
}

// The part that depends on the CTA size, then the launch.
static cudaError_t tiled_launch(TiledParams& p, bool select, int tq, cudaStream_t stream, bool sw = false)
{
    const Window& g = p.g;
    const size_t smem = tiled_smem(g, select, tq, &p.tile_cap, &p.tile_bytes);
    p.pitch = tq + g.kW;
    p.near_bound = sqrtf(g.d2max / 12.5f);
    if (select) {
        const int hh2 = g.kH / 2, hw2 = g.kW / 2;
        for (int j = 0; j < g.kt; ++j) {
            const int dh = p.walk[j] >> 16, dw = (int)(short)(p.walk[j] & 0xffff);
            p.walk_to[j] = ((dh + hh2) * p.pitch + (dw + hw2)) * 16;
        }
    }
    if (tq == 128) return launch_tiled_tq<128>(select, sw, p, smem, stream);
    if (tq == 160) return launch_tiled_tq<160>(select, sw, p, smem, stream);
    return launch_tiled_tq<192>(select, sw, p, smem, stream);
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
    if (force != 1 && total < (long long)dev.sm_count * 128 * 2) return 0;   // too few threads to fill the chip
    if (select && g.kt > MAX_WALK) return 0;

    TiledParams p;
    p.B = B; p.H = H; p.W = W; p.N = N; p.g = g;
    p.xyz1 = xyz1; p.xyz2 = xyz2; p.idx_n2 = idx_n2; p.random_hw = random_hw;
    p.out_nbr = nullptr; p.qs_ow = 1; p.qs_h = 1; p.qs_w = 1;
    p.out_idx = out_idx; p.out_valid = out_valid; p.out_vdis = out_vdis; p.out_mask = out_mask;
    p.total = total;
    tiled_prepare(p, select);

    // CTA size: the candidate whose CTAs all fit on the chip at once and load the SMs most evenly
    // CTAs per SM the register budget allows (tiled_min_ctas of the kernel template)
    // A store warp (one more warp per CTA that writes the count rows while the query warps walk) pays for itself
    // where the walk is long.  Measured on configs[0]'s frame (round 2, final kernels), without -> with: 11x41 139.2 ->
    // 139.4 us, 7x25 64.1 -> 62.9, 5x15 43.6 -> 43.9 (the query warps drop from 72 to 64 registers, and stores in
    // flight slow the walk's shared-memory loads even from another warp): used from 128 cells up.
    // elo_set_store_warp_min_cells / ELO_STORE_WARP_KT move the switch-over.
    const bool sw = select && (out_valid != nullptr || out_vdis != nullptr) &&
                    g.kt >= g_store_warp_min_kt.load(std::memory_order_relaxed);
    auto reg_ctas = [&](int tq) {
        if (!sw) return (select && g.K > 16 ? 512 : 896) / tq;
        if (g.K > 16) return 512 / (tq + 32);
        return tq == 128 ? 6 : tq == 160 ? 5 : 4;
    };
    int best_tq = 0;
    double best_cost = 0.0;
    for (int tq : {128, 160, 192}) {
        int cap, tb;
        const size_t smem = tiled_smem(g, select, tq, &cap, &tb);
        if (smem > 100 * 1024) continue;
        const long long ctas = (total + tq - 1) / tq;
        const long long per_sm = std::min<long long>((long long)((dev.max_smem_optin + 1024) / (smem + 1024)), (long long)reg_ctas(tq));
        if (per_sm < 1) continue;
        const long long rounds = (ctas + dev.sm_count - 1) / dev.sm_count;     // CTAs the busiest SM runs
        double cost = (double)rounds * tq;                                      // queries on the busiest SM
        if (ctas > per_sm * dev.sm_count) cost *= 1.25;                         // a second wave starts late
        if (best_tq == 0 || cost < best_cost) { best_tq = tq; best_cost = cost; }
    }
    if (best_tq == 0) return 0;
    const cudaError_t err = tiled_launch(p, select, best_tq, stream, sw);
    *rc = err == cudaSuccess ? ELO_OK : set_cuda_error(err, select ? "fused_conv_select_k (tiled) launch"
                                                                  : "fused_conv_random_k (tiled) launch");
    return 1;
}

// One search of elo_multi_search (dense query grid, compact (B, N, K) neighbour table out) on the tiled kernel:
// a quarter of the warp-per-query kernel's SM time for the wide windows, at a longer latency (128 queries per
// CTA, one thread each) -- what the throughput tile policy wants.  Returns 1 when it took the search.
int launch_search_tiled(bool select, int B, int H1, int W1, int oh, int ow, int qs_h, int qs_w, const Window& g,
                        const float* xyz1, const float* xyz2, const int* random_hw, int* out_nbr, cudaStream_t stream,
                        int* rc)
{
    if (g.K > 32 || !(g.d2max < 1e10f) || (select && g.kt > MAX_WALK)) return 0;
    int cap, tb;
    if (tiled_smem(g, select, 128, &cap, &tb) > 100 * 1024) return 0;
    TiledParams p;
    p.B = B; p.H = H1; p.W = W1; p.N = oh * ow; p.g = g;
    p.g.flag_copy = 0;
    p.xyz1 = xyz1; p.xyz2 = xyz2; p.idx_n2 = nullptr; p.random_hw = random_hw;
    p.out_nbr = out_nbr; p.qs_ow = ow; p.qs_h = qs_h; p.qs_w = qs_w;
    p.out_idx = nullptr; p.out_valid = nullptr; p.out_vdis = nullptr; p.out_mask = nullptr;
    p.total = (long long)B * p.N;
    tiled_prepare(p, select);
    const cudaError_t err = tiled_launch(p, select, 128, stream);
    *rc = err == cudaSuccess ? ELO_OK : set_cuda_error(err, "multi_search (tiled) launch");
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

extern "C" int elo_set_store_warp_min_cells(int min_cells)
{
    if (min_cells < 0) return elo::set_error(ELO_ERR_INVALID_ARGUMENT, "elo_set_store_warp_min_cells: >= 0");
    elo::g_store_warp_min_kt.store(min_cells, std::memory_order_relaxed);
    return ELO_OK;
}

extern "C" int elo_get_store_warp_min_cells(void) { return elo::g_store_warp_min_kt.load(std::memory_order_relaxed); }

extern "C" int elo_set_tile_staging(int mode)
{
    if (mode < 0 || mode > 1) return elo::set_error(ELO_ERR_INVALID_ARGUMENT, "elo_set_tile_staging: 0 or 1");
    elo::g_tile_staging.store(mode, std::memory_order_relaxed);
    return ELO_OK;
}

extern "C" int elo_get_tile_staging(void) { return elo::g_tile_staging.load(std::memory_order_relaxed); }
