// elo_tile.cuh -- pieces shared by the fused group kernels (set-conv, set-upconv, cost volume):
// query decoding, the in-tile neighbour search that fills the row -> cell table, and the gathers
// that build the first layer's input in shared memory.
#pragma once
#include "elo_mlp.cuh"
#include "elo_search.cuh"

namespace elo {

// The queries of a sample are an (oh x ow) sub-grid of the query image, cell (i*qs_h, j*qs_w):
// every pixel (qs = 1; utils/pointnet_util.py:23-30 get_hw_idx) or the strided centres a set-conv
// keeps (model_util.py:296-316 get_selected_idx).
struct QuerySet {
    int H1, W1;      // query image
    int oh, ow;      // queries per sample = oh * ow
    int qs_h, qs_w;  // query stride inside the query image
};

__device__ __forceinline__ void query_cell(const QuerySet& qs, int n, int& h, int& w)
{
    h = (n / qs.ow) * qs.qs_h;
    w = (n % qs.ow) * qs.qs_w;
}

// Shared-memory carve-up helper (all regions 16 B aligned).
struct SmemCarver {
    unsigned char* p;
    __device__ explicit SmemCarver(unsigned char* base) : p(base) {}
    template <typename T>
    __device__ T* take(size_t count)
    {
        T* r = reinterpret_cast<T*>(p);
        p += (count * sizeof(T) + 15) & ~size_t(15);
        return r;
    }
};

// Neighbour search for the tile's queries, one warp per query.
//   nbr[q * K + k] = linear cell (hh * w2 + ww) of the k-th neighbour in the searched grid, -1 if masked
//   ctr[q * 4 + {0,1,2}] = centre xyz,  ctr[q*4+3] = batch index as int bits (-1: padding query)
// SELECT needs per-warp scratch (dist, hw) of kt entries each.
template <bool SELECT>
__device__ __forceinline__ void tile_search(const QuerySet& qs, const Window& g, const float* __restrict__ xyz1,
                                            const float* __restrict__ xyz2, const int2* off, long long q0,
                                            int qt, long long total_q, int* nbr, float* ctr, float* scratch_dist,
                                            int* scratch_hw)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = COMPUTE_WARPS;
    const int nq = qs.oh * qs.ow;
    for (int ql = warp; ql < qt; ql += nwarps) {
        int* row = nbr + ql * g.K;
        for (int k = lane; k < g.K; k += 32) row[k] = -1;
        const long long gq = q0 + ql;
        float xc = 0.f, yc = 0.f, zc = 0.f;
        int b = -1, h = 0, w = 0;
        if (gq < total_q) {
            b = (int)(gq / nq);
            query_cell(qs, (int)(gq % nq), h, w);
            const float* c = xyz1 + ((size_t)b * qs.H1 * qs.W1 + (size_t)h * qs.W1 + w) * 3;
            xc = __ldg(c); yc = __ldg(c + 1); zc = __ldg(c + 2);
        }
        if (lane == 0) {
            ctr[ql * 4 + 0] = xc; ctr[ql * 4 + 1] = yc; ctr[ql * 4 + 2] = zc;
            ctr[ql * 4 + 3] = __int_as_float(b);
        }
        __syncwarp();
        if (b < 0 || fmaxf(sq3(xc, yc, zc), 1e-10f) <= 1e-10f) continue;   // padding / empty centre: all masked
        const float* g2 = xyz2 + (size_t)b * g.h2 * g.w2 * 3;
        const int ch = h / g.stride_h, cw = w / g.stride_w;
        auto emit = [&](int slot, int hh, int ww) { row[slot] = hh * g.w2 + ww; };
        if (SELECT) {
            int written;
            search_select_k(g2, off, g, ch, cw, xc, yc, zc, scratch_dist + (size_t)warp * g.kt,
                            scratch_hw + (size_t)warp * g.kt, &written, emit);
        } else {
            search_random_k(g2, off, g, ch, cw, xc, yc, zc, emit);
        }
        __syncwarp();
    }
}

// Same tables as tile_search, but read from a pre-computed nbr table (elo_multi_search) instead of searching.
__device__ __forceinline__ void tile_load_nbr(const QuerySet& qs, int K, const float* __restrict__ xyz1,
                                              const int* __restrict__ nbr_in, long long q0, int qt,
                                              long long total_q, int* nbr, float* ctr)
{
    const int nq = qs.oh * qs.ow;
    for (int t = threadIdx.x; t < qt * K; t += CTA_THREADS) {
        const long long gq = q0 + t / K;
        nbr[t] = gq < total_q ? __ldg(nbr_in + q0 * K + t) : -1;
    }
    for (int ql = threadIdx.x; ql < qt; ql += CTA_THREADS) {
        const long long gq = q0 + ql;
        float xc = 0.f, yc = 0.f, zc = 0.f;
        int b = -1;
        if (gq < total_q) {
            int h, w;
            b = (int)(gq / nq);
            query_cell(qs, (int)(gq % nq), h, w);
            const float* c = xyz1 + ((size_t)b * qs.H1 * qs.W1 + (size_t)h * qs.W1 + w) * 3;
            xc = __ldg(c); yc = __ldg(c + 1); zc = __ldg(c + 2);
        }
        ctr[ql * 4 + 0] = xc; ctr[ql * 4 + 1] = yc; ctr[ql * 4 + 2] = zc;
        ctr[ql * 4 + 3] = __int_as_float(b);
    }
}

// X[c0 + c][r] = src[(b, cell(r)), c] for c < C (C % 4 == 0), zero for masked / padding rows.
// cell_of(r) returns the linear row of `src` (>= 0) or -1.
template <int U = 4, typename CellOf>
__device__ __forceinline__ void gather_features(float* X, int RS, int c0, const float* __restrict__ src, int C,
                                                int rows, CellOf cell_of)
{
    // U tasks per thread and trip: their (L2-latency) loads are in flight together
    const int c4n = C >> 2, total = rows * c4n;
    for (int t0 = threadIdx.x; t0 < total; t0 += CTA_THREADS * U) {
        float4 v[U];
        int rr[U], cc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u * CTA_THREADS;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            rr[u] = -1; cc[u] = 0;
            if (t < total) {
                const int r = t / c4n, c4 = t - r * c4n;
                const long long cell = cell_of(r);
                if (cell >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)cell * C) + c4);
                rr[u] = r; cc[u] = c0 + c4 * 4;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (rr[u] < 0) continue;
            X[act_index(cc[u] + 0, rr[u], RS)] = v[u].x;
            X[act_index(cc[u] + 1, rr[u], RS)] = v[u].y;
            X[act_index(cc[u] + 2, rr[u], RS)] = v[u].z;
            X[act_index(cc[u] + 3, rr[u], RS)] = v[u].w;
        }
    }
}

// exact (non-contracted) |d|^2 as TensorFlow's square + reduce_sum would compute it
__device__ __forceinline__ float sumsq_tf(float dx, float dy, float dz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace elo
