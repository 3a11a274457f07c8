// fused_blocks.cu -- the neighbour-search -> gather -> shared-MLP -> pool blocks of the network, each
// as ONE kernel whose (B,N,K,C) intermediates never reach HBM (sm_100a).
//
//   elo_group_mlp_max   set-conv / first half of set-upconv   utils/pointnet_util.py:197-230, 272-298
//   elo_cost_volume_1   cost volume, point-to-patch stage      utils/pointnet_util.py:45-100
//   elo_cost_volume_2   cost volume, patch-to-patch stage      utils/pointnet_util.py:104-146
//   elo_row_mlp         per-point MLP chains                   utils/pointnet_util.py:153-175, 303-311
//
// Every kernel: (1) one thread starts the weight stream (TMA bulk copies, elo_mlp.cuh) so the first
// chunks land while (2) warps run the projection-aware neighbour search of the tile's queries
// (elo_search.cuh, same code as the stand-alone index ops) and (3) all threads gather xyz / features
// of the selected cells into the swizzled [channel][row] tile; (4) the layers run back to back out
// of shared memory; (5) the reduction over the K neighbours (max, or masked softmax-weighted sum)
// writes (B,N,C_out) once.  Inference batch-norm is folded into the weights on the host.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/elo_b200.h"
#include "elo_common.cuh"
#include "elo_tc_engine.cuh"
#include "elo_tile.cuh"

namespace elo {

static constexpr int SMEM_LIMIT = 227 * 1024;

// ================================================================================================
// group_mlp_max
struct GroupMlpParams;
__device__ __forceinline__ bool g_direct_gather_dev(const GroupMlpParams& p);

struct GroupMlpParams {
    int direct_gather;            // tensor-core kernel: gather straight into tensor memory (A/B switch, default on)
    QuerySet qs;
    Window g;
    long long q_base[2], q_end[2];   // global query range of each parameter set
    int qt, Cf, nl, cout[3], total_chunks, nring;
    const float* xyz1;
    const float* xyz2;
    const float* feat2[2];
    const int* random_hw[2];
    const float* weights[2];
    float* out[2];
    int* dbg_nbr[2];
    const int* nbr_in[2];
    long long* tlog;
};

template <int NB>
__global__ void __launch_bounds__(LAUNCH_THREADS, 1) group_mlp_max_kernel(const GroupMlpParams p)
{
    constexpr int RS = NB * 64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();          // cross-check engine: no prologue overlap, plain stream order from here
    SmemCarver sc(smem_raw);
    float* ring = sc.take<float>((size_t)p.nring * CHUNK_FLOATS);
    uint64_t* bars = sc.take<uint64_t>(2 * MAX_RING);
    int2* off = sc.take<int2>(p.g.kt);
    int* nbr = sc.take<int>(RS);
    float* ctr = sc.take<float>(RS * 4);
    const int cin0 = 3 + p.Cf;
    float* X = sc.take<float>((size_t)((cin0 + 3) & ~3) * RS);
    float* A = sc.take<float>(128 * RS);
    float* Bf = sc.take<float>(128 * RS);

    const int set = blockIdx.y;
    const Window g = p.g;
    WeightStream ws;
    ws.init(ring, bars, p.nring);
    if (threadIdx.x >= CTA_THREADS) {            // producer warp: nothing but the weight stream
        ws.produce(p.weights[set], p.total_chunks, 1);
        return;
    }
    build_offsets(off, p.random_hw[set], g.kt, g.kH, g.kW, CTA_THREADS);
    for (int i = threadIdx.x; i < RS; i += CTA_THREADS) nbr[i] = -1;
    compute_sync();

    const long long q0 = p.q_base[set] + (long long)blockIdx.x * p.qt;
    const long long total_q = p.q_end[set];
    const int cells2 = g.h2 * g.w2;
    if (p.nbr_in[set] != nullptr) tile_load_nbr(p.qs, g.K, p.xyz1, p.nbr_in[set], q0, p.qt, total_q, nbr, ctr);
    else tile_search<false>(p.qs, g, p.xyz1, p.xyz2, off, q0, p.qt, total_q, nbr, ctr, nullptr, nullptr);
    compute_sync();

    // first layer input: [q_k - p (3), feat2_k (Cf)], masked neighbours contribute q = 0, feat = 0
    for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
        const int q = r / g.K;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (q < p.qt) {
            const int b = __float_as_int(ctr[q * 4 + 3]);
            if (b >= 0) {
                float qx = 0.f, qy = 0.f, qz = 0.f;
                const int cell = nbr[r];
                if (cell >= 0) {
                    const float* s = p.xyz2 + ((size_t)b * cells2 + cell) * 3;
                    qx = __ldg(s); qy = __ldg(s + 1); qz = __ldg(s + 2);
                }
                dx = qx - ctr[q * 4 + 0]; dy = qy - ctr[q * 4 + 1]; dz = qz - ctr[q * 4 + 2];
            }
        }
        X[act_index(0, r, RS)] = dx;
        X[act_index(1, r, RS)] = dy;
        X[act_index(2, r, RS)] = dz;
    }
    {
        const float* feat = p.feat2[set];
        const int K = g.K, qt = p.qt;
        gather_features(X, RS, 3, feat, p.Cf, RS, [&](int r) -> long long {
            const int q = r / K;
            if (q >= qt || nbr[r] < 0) return -1;
            return (long long)__float_as_int(ctr[q * 4 + 3]) * cells2 + nbr[r];
        });
    }
    compute_sync();

    const float* in = X;
    int cin = cin0;
    for (int l = 0; l < p.nl; ++l) {
        float* o = (l & 1) ? Bf : A;
        dense_rt<NB>(ws, in, cin, o, p.cout[l]);
        in = o;
        cin = p.cout[l];
    }

    // max over the K neighbours of (y * mask): y >= 0 after ReLU, masked rows count as 0
    const int Cout = cin;
    float* out = p.out[set];
    for (int t = threadIdx.x; t < p.qt * Cout; t += CTA_THREADS) {
        const int q = t / Cout, c = t - q * Cout;
        const long long gq = q0 + q;
        if (gq >= total_q) break;
        float m = 0.f;
        for (int k = 0; k < g.K; ++k)
            if (nbr[q * g.K + k] >= 0) m = fmaxf(m, in[act_index(c, q * g.K + k, RS)]);
        out[gq * Cout + c] = m;
    }
    if (p.dbg_nbr[set] != nullptr)
        for (int t = threadIdx.x; t < p.qt * g.K; t += CTA_THREADS)
            if (q0 + t / g.K < total_q) p.dbg_nbr[set][q0 * g.K + t] = nbr[t];
}

// ================================================================================================
// cost volume, stage 1 (point-to-patch)
struct Cv1Params {
    int direct_gather;            // tensor-core kernel: gather straight into tensor memory (A/B switch, default on)
    QuerySet qs;
    Window g;
    long long q_first, total_q;      // pixels [q_first, total_q) are this call's queries
    int qt, C, total_chunks, nring;
    int sl_in_ring;      // tensor-core engine: the logits' pool staging aliases the (by then dead) weight ring
    const float* xyz1;
    const float* xyz2;
    const float* f1;
    const float* f2;
    const int* random_hw;
    const float* weights;
    float* out;
    int* dbg_nbr;
    const int* nbr_in;
    long long* tlog;
};

// xyz part of a cost-volume row: [p, q, q - p, sqrt(|q - p|^2 + 1e-20)] (utils/pointnet_util.py:60-63)
__device__ __forceinline__ void write_xyz10(float* X, int RS, int r, float px, float py, float pz, float qx,
                                            float qy, float qz)
{
    const float dx = qx - px, dy = qy - py, dz = qz - pz;
    X[act_index(0, r, RS)] = px; X[act_index(1, r, RS)] = py; X[act_index(2, r, RS)] = pz;
    X[act_index(3, r, RS)] = qx; X[act_index(4, r, RS)] = qy; X[act_index(5, r, RS)] = qz;
    X[act_index(6, r, RS)] = dx; X[act_index(7, r, RS)] = dy; X[act_index(8, r, RS)] = dz;
    X[act_index(9, r, RS)] = sqrtf(__fadd_rn(sumsq_tf(dx, dy, dz), 1e-20f));
}

// masked softmax over the K rows of a query, per channel, applied to `val` (TF: where(mask, w, -1e10),
// softmax(dim=2), reduce_sum(w * val)).  An all-masked group gets uniform weights 1/K.
__device__ __forceinline__ float softmax_pool(const float* logit, const float* val, const int* nbr_row, int c,
                                              int row0, int K, int RS)
{
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
        const float l = nbr_row[k] >= 0 ? logit[act_index(c, row0 + k, RS)] : -1e10f;
        m = fmaxf(m, l);
    }
    float s = 0.f, acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const float l = nbr_row[k] >= 0 ? logit[act_index(c, row0 + k, RS)] : -1e10f;
        const float e = expf(l - m);
        s += e;
        acc = fmaf(e, val[act_index(c, row0 + k, RS)], acc);
    }
    return acc / s;
}

template <int NB>
__global__ void __launch_bounds__(LAUNCH_THREADS, 1) cost_volume_1_kernel(const Cv1Params p)
{
    constexpr int RS = NB * 64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();          // cross-check engine: no prologue overlap, plain stream order from here
    SmemCarver sc(smem_raw);
    float* ring = sc.take<float>((size_t)p.nring * CHUNK_FLOATS);
    uint64_t* bars = sc.take<uint64_t>(2 * MAX_RING);
    int2* off = sc.take<int2>(p.g.kt);
    int* nbr = sc.take<int>(RS);
    float* ctr = sc.take<float>(RS * 4);
    const int C = p.C, xc = 10 + 2 * C;
    float* X = sc.take<float>((size_t)((xc + 3) & ~3) * RS);
    float* A = sc.take<float>(128 * RS);     // A, Bf, Cb are contiguous: the select-K scratch aliases them
    float* Bf = sc.take<float>(128 * RS);
    float* Cb = sc.take<float>(64 * RS);

    const Window g = p.g;
    WeightStream ws;
    ws.init(ring, bars, p.nring);
    if (threadIdx.x >= CTA_THREADS) {            // producer warp: nothing but the weight stream
        ws.produce(p.weights, p.total_chunks, 1);
        return;
    }
    build_offsets(off, p.random_hw, g.kt, g.kH, g.kW, CTA_THREADS);
    for (int i = threadIdx.x; i < RS; i += CTA_THREADS) nbr[i] = -1;
    compute_sync();

    const long long q0 = p.q_first + (long long)blockIdx.x * p.qt;
    const int cells = g.h2 * g.w2, nwarps = COMPUTE_WARPS;
    float* sdist = A;
    int* shw = reinterpret_cast<int*>(A + (size_t)nwarps * g.kt);
    if (p.nbr_in != nullptr) tile_load_nbr(p.qs, g.K, p.xyz1, p.nbr_in, q0, p.qt, p.total_q, nbr, ctr);
    else tile_search<true>(p.qs, g, p.xyz1, p.xyz2, off, q0, p.qt, p.total_q, nbr, ctr, sdist, shw);
    compute_sync();

    for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
        const int q = r / g.K;
        float px = 0.f, py = 0.f, pz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
        if (q < p.qt) {
            const int b = __float_as_int(ctr[q * 4 + 3]);
            if (b >= 0) {
                px = ctr[q * 4 + 0]; py = ctr[q * 4 + 1]; pz = ctr[q * 4 + 2];
                const int cell = nbr[r];
                if (cell >= 0) {
                    const float* s = p.xyz2 + ((size_t)b * cells + cell) * 3;
                    qx = __ldg(s); qy = __ldg(s + 1); qz = __ldg(s + 2);
                }
            }
        }
        write_xyz10(X, RS, r, px, py, pz, qx, qy, qz);
    }
    {
        const int K = g.K, qt = p.qt, nq = p.qs.oh * p.qs.ow;
        const long long total = p.total_q;
        // f1 of the query pixel, repeated for its K rows (queries are all pixels: linear cell = gq)
        gather_features(X, RS, 10, p.f1, C, RS, [&](int r) -> long long {
            const int q = r / K;
            return (q < qt && q0 + q < total) ? q0 + q : -1;
        });
        gather_features(X, RS, 10 + C, p.f2, C, RS, [&](int r) -> long long {
            const int q = r / K;
            if (q >= qt || nbr[r] < 0) return -1;
            return (long long)((q0 + q) / nq) * cells + nbr[r];
        });
    }
    compute_sync();

    dense<NB, 128, true>(ws, X, xc, A);            // CV_0
    dense<NB, 64, true>(ws, A, 128, Cb);           // CV_1
    dense<NB, 64, true>(ws, Cb, 64, Bf + 64 * RS); // CV_2      -> F   = Bf[64:128]
    dense<NB, 64, true>(ws, X, 10, Bf);            // CV_xyz    -> enc = Bf[0:64]
    dense<NB, 128, true>(ws, Bf, 128, A);          // sum_CV_0 on [enc, F]
    dense<NB, 64, true>(ws, A, 128, Cb);           // sum_CV_1  -> attention logits

    for (int t = threadIdx.x; t < p.qt * 64; t += CTA_THREADS) {
        const int q = t >> 6, c = t & 63;
        const long long gq = q0 + q;
        if (gq >= p.total_q) break;
        p.out[gq * 64 + c] = softmax_pool(Cb, Bf + 64 * RS, nbr + q * g.K, c, q * g.K, g.K, RS);
    }
    if (p.dbg_nbr != nullptr)
        for (int t = threadIdx.x; t < p.qt * g.K; t += CTA_THREADS)
            if (q0 + t / g.K < p.total_q) p.dbg_nbr[q0 * g.K + t] = nbr[t];
}

// ================================================================================================
// cost volume, stage 2 (patch-to-patch)
struct Cv2Params {
    QuerySet qs;
    Window g;
    long long q_first, total_q;      // pixels [q_first, total_q) are this call's queries
    int qt, C, total_chunks, nring;
    int sl_in_ring;      // tensor-core engine: logits staged in the dead weight ring, values pooled straight from X
    const float* xyz1;
    const float* f1;
    const float* cv1;
    const int* random_hw;
    const float* weights;
    float* out;
    int* dbg_nbr;
    const int* nbr_in;
    long long* tlog;
};

template <int NB>
__global__ void __launch_bounds__(LAUNCH_THREADS, 1) cost_volume_2_kernel(const Cv2Params p)
{
    constexpr int RS = NB * 64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();          // cross-check engine: no prologue overlap, plain stream order from here
    SmemCarver sc(smem_raw);
    float* ring = sc.take<float>((size_t)p.nring * CHUNK_FLOATS);
    uint64_t* bars = sc.take<uint64_t>(2 * MAX_RING);
    int2* off = sc.take<int2>(p.g.kt);
    int* nbr = sc.take<int>(RS);
    float* ctr = sc.take<float>(RS * 4);
    const int C = p.C;
    float* Y = sc.take<float>(12 * RS);
    float* X = sc.take<float>((size_t)(128 + C) * RS);
    float* A = sc.take<float>(128 * RS);
    float* Cb = sc.take<float>(64 * RS);

    const Window g = p.g;
    WeightStream ws;
    ws.init(ring, bars, p.nring);
    if (threadIdx.x >= CTA_THREADS) {            // producer warp: nothing but the weight stream
        ws.produce(p.weights, p.total_chunks, 1);
        return;
    }
    build_offsets(off, p.random_hw, g.kt, g.kH, g.kW, CTA_THREADS);
    for (int i = threadIdx.x; i < RS; i += CTA_THREADS) nbr[i] = -1;
    compute_sync();

    const long long q0 = p.q_first + (long long)blockIdx.x * p.qt;
    const int cells = g.h2 * g.w2;
    if (p.nbr_in != nullptr) tile_load_nbr(p.qs, g.K, p.xyz1, p.nbr_in, q0, p.qt, p.total_q, nbr, ctr);
    else tile_search<false>(p.qs, g, p.xyz1, p.xyz1, off, q0, p.qt, p.total_q, nbr, ctr, nullptr, nullptr);
    compute_sync();

    for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
        const int q = r / g.K;
        float px = 0.f, py = 0.f, pz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
        if (q < p.qt) {
            const int b = __float_as_int(ctr[q * 4 + 3]);
            if (b >= 0) {
                px = ctr[q * 4 + 0]; py = ctr[q * 4 + 1]; pz = ctr[q * 4 + 2];
                const int cell = nbr[r];
                if (cell >= 0) {
                    const float* s = p.xyz1 + ((size_t)b * cells + cell) * 3;
                    qx = __ldg(s); qy = __ldg(s + 1); qz = __ldg(s + 2);
                }
            }
        }
        write_xyz10(Y, RS, r, px, py, pz, qx, qy, qz);
    }
    {
        const int K = g.K, qt = p.qt, nq = p.qs.oh * p.qs.ow;
        const long long total = p.total_q;
        gather_features(X, RS, 64, p.f1, C, RS, [&](int r) -> long long {
            const int q = r / K;
            return (q < qt && q0 + q < total) ? q0 + q : -1;
        });
        gather_features(X, RS, 64 + C, p.cv1, 64, RS, [&](int r) -> long long {
            const int q = r / K;
            if (q >= qt || nbr[r] < 0) return -1;
            return (long long)((q0 + q) / nq) * cells + nbr[r];
        });
    }
    compute_sync();

    dense<NB, 64, true>(ws, Y, 10, X);              // sum_xyz_encoding -> X[0:64]
    dense<NB, 128, true>(ws, X, 128 + C, A);        // sum_cost_volume_0 on [enc, f1, stage-1 of neighbour]
    dense<NB, 64, true>(ws, A, 128, Cb);            // sum_cost_volume_1 -> attention logits

    // weights applied to the gathered stage-1 features, which sit at channel 64 + C (any alignment:
    // softmax_pool indexes through act_index with absolute channels)
    for (int t = threadIdx.x; t < p.qt * 64; t += CTA_THREADS) {
        const int q = t >> 6, c = t & 63;
        const long long gq = q0 + q;
        if (gq >= p.total_q) break;
        const int* nrow = nbr + q * g.K;
        float m = -INFINITY;
        for (int k = 0; k < g.K; ++k)
            m = fmaxf(m, nrow[k] >= 0 ? Cb[act_index(c, q * g.K + k, RS)] : -1e10f);
        float s = 0.f, acc = 0.f;
        for (int k = 0; k < g.K; ++k) {
            const float l = nrow[k] >= 0 ? Cb[act_index(c, q * g.K + k, RS)] : -1e10f;
            const float e = expf(l - m);
            s += e;
            acc = fmaf(e, X[act_index(64 + C + c, q * g.K + k, RS)], acc);
        }
        p.out[gq * 64 + c] = acc / s;
    }
    if (p.dbg_nbr != nullptr)
        for (int t = threadIdx.x; t < p.qt * g.K; t += CTA_THREADS)
            if (q0 + t / g.K < p.total_q) p.dbg_nbr[q0 * g.K + t] = nbr[t];
}

// ================================================================================================
// row_mlp: per-point MLP chains on concatenated (rows, C_i) tensors; up to two phases, the second
// may take the first's output as one of its sources (set-upconv's second half feeding a predictor).
struct RowMlpParams {
    long long rows;
    int rt;                    // rows per tile
    int nphase;
    int nsrc[2];
    int src_c[2][3];           // channels of each source
    int src_prev[2][3];        // 1: the source is the previous phase's output (kept in shared memory)
    int nl[2], cout[2][3];
    int total_chunks, nring, stage_ch;
    int slot[2][3], stage_here[2][3];   // tensor-core path: staging channel of each source / stage it here
    const float* src[2][2][3]; // [set][phase][i]
    const float* weights[2];
    float* out[2];
    float* out_phase0[2];      // optional: also store phase 0's result (rows, cout) -- may be null
    long long* tlog;
};

template <int NB>
__global__ void __launch_bounds__(LAUNCH_THREADS, 1) row_mlp_kernel(const RowMlpParams p)
{
    constexpr int RS = NB * 64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();          // cross-check engine: no prologue overlap, plain stream order from here
    SmemCarver sc(smem_raw);
    float* ring = sc.take<float>((size_t)p.nring * CHUNK_FLOATS);
    uint64_t* bars = sc.take<uint64_t>(2 * MAX_RING);
    int xmax = 0;
    for (int ph = 0; ph < p.nphase; ++ph) {
        int c = 0;
        for (int i = 0; i < p.nsrc[ph]; ++i) c += p.src_c[ph][i];
        xmax = max(xmax, c);
    }
    float* X = sc.take<float>((size_t)((xmax + 3) & ~3) * RS);
    float* A = sc.take<float>(128 * RS);
    float* Bf = sc.take<float>(128 * RS);

    const int set = blockIdx.y;
    WeightStream ws;
    ws.init(ring, bars, p.nring);
    if (threadIdx.x >= CTA_THREADS) {
        ws.produce(p.weights[set], p.total_chunks, 1);
        return;
    }
    const long long r0 = (long long)blockIdx.x * p.rt;
    const long long rows = p.rows;
    const int rt = p.rt;

    const float* prev = nullptr;
    int prev_c = 0;
    for (int ph = 0; ph < p.nphase; ++ph) {
        int c0 = 0;
        for (int i = 0; i < p.nsrc[ph]; ++i) {
            const int Ci = p.src_c[ph][i];
            if (p.src_prev[ph][i]) {
                for (int t = threadIdx.x; t < RS * prev_c; t += CTA_THREADS) {
                    const int c = t / RS, r = t - c * RS;
                    X[act_index(c0 + c, r, RS)] = prev[act_index(c, r, RS)];
                }
            } else {
                gather_features(X, RS, c0, p.src[set][ph][i], Ci, RS, [&](int r) -> long long {
                    return (r < rt && r0 + r < rows) ? r0 + r : -1;
                });
            }
            c0 += Ci;
        }
        compute_sync();
        const float* in = X;
        int cin = c0;
        for (int l = 0; l < p.nl[ph]; ++l) {
            float* o = (l & 1) ? Bf : A;
            dense_rt<NB>(ws, in, cin, o, p.cout[ph][l]);
            in = o;
            cin = p.cout[ph][l];
        }
        prev = in;
        prev_c = cin;
        float* dst = (ph == p.nphase - 1) ? p.out[set] : p.out_phase0[set];
        if (dst != nullptr)
            for (int t = threadIdx.x; t < rt * cin; t += CTA_THREADS) {
                const int r = t / cin, c = t - r * cin;
                if (r0 + r >= rows) break;
                dst[(r0 + r) * cin + c] = in[act_index(c, r, RS)];
            }
        // the next phase's gather overwrites X only; `prev` (A or Bf) stays valid because the
        // first layer of the next phase writes A only after reading X... unless prev == A:
        // copy-out above happens before any write, and the prev->X copy is fenced by the
        // __syncthreads() before the dense chain.
    }
}

// ================================================================================================
// Tensor-core (tcgen05, 3xTF32) versions of the four kernels.  Same tiles, same search / gather /
// pooling code; the layers run on the TcPipe engine (elo_tc_engine.cuh): activations in TMEM, weights
// streamed by the TMA warp, MMAs issued by a dedicated warp.  Tiles have TC_ROWS = 128 rows.
// `weights` here = [TC stream: total_chunks x 4096 floats][biases of all layers, in order].

constexpr int TC_MAX_BIAS = 768;     // channels summed over the layers of one kernel (cost volume stage 1: 512)

struct TcSmem {
    float* ring; uint64_t* bars; uint32_t* tmem_holder; float* bias; int2* off; int* nbr; float* ctr; float* X;
};

__device__ __forceinline__ TcSmem tc_carve(unsigned char* raw, int nring, int kt, size_t staging_floats)
{
    SmemCarver sc(raw);
    TcSmem m;
    m.ring = sc.take<float>((size_t)nring * TC_CHUNK_FLOATS);
    m.bars = sc.take<uint64_t>(2 * MAX_RING + 2);
    m.tmem_holder = sc.take<uint32_t>(4);
    m.bias = sc.take<float>(TC_MAX_BIAS);
    m.off = sc.take<int2>(kt);
    m.nbr = sc.take<int>(TC_ROWS);
    m.ctr = sc.take<float>(TC_ROWS * 4);
    m.X = sc.take<float>(staging_floats);
    return m;
}

// Row layout of a tensor-core group tile: the K rows of a query never straddle a 32-row lane quarter, so
// the reduction over the K neighbours is a warp shuffle among the K lanes that own them.
//   gpw = 32 / K queries per quarter;  query ql -> rows (ql / gpw) * 32 + (ql % gpw) * K + k,  k < K.
struct TcRows {
    int K, gpw;
    __device__ __forceinline__ explicit TcRows(int K_) : K(K_), gpw(32 / K_) {}
    __device__ __forceinline__ int row(int ql, int k) const { return (ql / gpw) * 32 + (ql % gpw) * K + k; }
    // query / neighbour of a row, or ql = -1 for the padding rows at the end of a quarter
    __device__ __forceinline__ void decode(int r, int& ql, int& k) const
    {
        const int w = r & 31;
        if (w >= gpw * K) { ql = -1; k = 0; return; }
        ql = (r >> 5) * gpw + w / K;
        k = w % K;
    }
};

// nbr[row] / ctr[ql] of the tile from a pre-computed neighbour table (rows of padding / absent queries: -1)
__device__ __forceinline__ void tc_load_nbr(const QuerySet& qs, const TcRows& rows, const float* __restrict__ xyz1,
                                            const int* __restrict__ nbr_in, long long q0, int qt, long long total_q,
                                            int* nbr, float* ctr)
{
    const int nq = qs.oh * qs.ow;
    for (int r = threadIdx.x; r < TC_ROWS; r += CTA_THREADS) {
        int ql, k;
        rows.decode(r, ql, k);
        nbr[r] = (ql >= 0 && ql < qt && q0 + ql < total_q) ? __ldg(nbr_in + (q0 + ql) * rows.K + k) : -1;
    }
    for (int ql = threadIdx.x; ql < TC_ROWS; ql += CTA_THREADS) {
        const long long gq = q0 + ql;
        float xc = 0.f, yc = 0.f, zc = 0.f;
        int b = -1;
        if (ql < qt && gq < total_q) {
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

// centre / neighbour xyz of a row (zeros for masked neighbours and padding rows)
__device__ __forceinline__ void tc_row_xyz(const TcRows& rows, int r, int qt, const int* nbr, const float* ctr,
                                           const float* __restrict__ xyz2, int cells2, float (&p)[3], float (&q)[3])
{
    int ql, k;
    rows.decode(r, ql, k);
    p[0] = p[1] = p[2] = q[0] = q[1] = q[2] = 0.f;
    if (ql < 0 || ql >= qt) return;
    const int b = __float_as_int(ctr[ql * 4 + 3]);
    if (b < 0) return;
    p[0] = ctr[ql * 4 + 0]; p[1] = ctr[ql * 4 + 1]; p[2] = ctr[ql * 4 + 2];
    const int cell = nbr[r];
    if (cell >= 0) {
        const float* s = xyz2 + ((size_t)b * cells2 + cell) * 3;
        q[0] = __ldg(s); q[1] = __ldg(s + 1); q[2] = __ldg(s + 2);
    }
}

// Pool staging: S[row][POOL_LD] with POOL_LD odd -> conflict-free both for the epilogue (lanes = consecutive
// rows, one channel) and for the pooling tasks (lanes = consecutive channels, one row).
constexpr int POOL_LD64 = 65;

// masked softmax over the K rows of a query for one channel, applied to val (TF: where(mask, w, -1e10),
// softmax(dim=2), reduce_sum(w * val)); an all-masked group gets uniform weights 1/K
template <typename ValueAt>
__device__ __forceinline__ float pool_softmax_v(const float* SL, ValueAt value_at, const int* nbr, int r0, int K, int c)
{
    // Online form over chunks of 8 rows: a chunk's loads, its exponentials and its sums are independent of each
    // other (K = 4 / 6 / 32 in the model: one chunk is the common case), only the running maximum couples the
    // chunks.  __expf (ex2.approx) is accurate to ~1e-6 relative on the arguments (<= 0) seen here, two orders
    // below the parity bar.
    float m = -INFINITY, s = 0.f, a = 0.f;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float l[8], v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            l[i] = -INFINITY; v[i] = 0.f;
            if (k < K) {
                l[i] = nbr[r0 + k] >= 0 ? SL[(r0 + k) * POOL_LD64 + c] : -1e10f;
                v[i] = value_at(r0 + k, c);
            }
        }
        float mn = m;
#pragma unroll
        for (int i = 0; i < 8; ++i) mn = fmaxf(mn, l[i]);
        const float sc = __expf(m - mn);          // first chunk: exp(-inf) = 0 on s = a = 0
        s *= sc; a *= sc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float e = __expf(l[i] - mn);      // rows past K: exp(-inf) = 0
            s += e;
            a = fmaf(e, v[i], a);
        }
        m = mn;
    }
    return a / s;
}

__device__ __forceinline__ float pool_softmax(const float* SL, const float* SV, int ld_v, const int* nbr, int r0, int K,
                                              int c)
{
    return pool_softmax_v(SL, [&](int row, int ch) { return SV[row * ld_v + ch]; }, nbr, r0, K, c);
}

__device__ __forceinline__ bool g_direct_gather_dev(const GroupMlpParams& p) { return p.direct_gather != 0; }

__global__ void __launch_bounds__(TC_LAUNCH_THREADS, 2) group_mlp_max_tc_kernel(const GroupMlpParams p)
{
    constexpr int RS = TC_ROWS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cin0 = 3 + p.Cf;
    const int cout_last = p.cout[p.nl - 1];
    const int pool_ld = cout_last + 1;
    TcSmem sm = tc_carve(smem_raw, p.nring, p.g.kt, (size_t)max(((cin0 + 3) & ~3), pool_ld) * RS);
    const int set = blockIdx.y, warp = threadIdx.x >> 5;
    const Window g = p.g;
    const TcRows rows(g.K);
    int* nbr = sm.nbr; float* ctr = sm.ctr; float* X = sm.X;
    const long long q0 = p.q_base[set] + (long long)blockIdx.x * p.qt;
    const long long total_q = p.q_end[set];
    const int cells2 = g.h2 * g.w2;
    // Prologue on constants only (barriers, tensor memory, biases; the producer warp starts streaming weights):
    // under programmatic dependent launch it runs while the kernel before this one is still finishing.
    pdl_trigger();
    TcPipe pipe;
    int nbias = 0;
    for (int l = 0; l < p.nl; ++l) nbias += p.cout[l];
    pipe.begin(sm.ring, sm.bars, p.nring, sm.tmem_holder, p.weights[set] + (size_t)p.total_chunks * TC_CHUNK_FLOATS,
               sm.bias, nbias, p.tlog, /*alloc_now=*/false);
    if (warp < COMPUTE_WARPS) {
        pdl_wait();                              // from here on: data written by earlier kernels
        if (p.nbr_in[set] != nullptr) tc_load_nbr(p.qs, rows, p.xyz1, p.nbr_in[set], q0, p.qt, total_q, nbr, ctr);
    }
    pipe.join();
    if (warp == COMPUTE_WARPS) { pipe.produce(p.weights[set], p.total_chunks); return; }
    if (warp == COMPUTE_WARPS + 1) {
        if (tc::elect_one()) {
            int cin = cin0;
            for (int l = 0; l < p.nl; ++l) { pipe.issue_layer(0, cin, p.cout[l]); cin = p.cout[l]; }
        }
        return;
    }
    if (p.nbr_in[set] == nullptr) {
        // stand-alone call without a table: search here, into a contiguous scratch, then re-lay the rows
        build_offsets(sm.off, p.random_hw[set], g.kt, g.kH, g.kW, CTA_THREADS);
        int* tmp = reinterpret_cast<int*>(X);
        for (int i = threadIdx.x; i < RS; i += CTA_THREADS) tmp[i] = -1;
        compute_sync();
        tile_search<false>(p.qs, g, p.xyz1, p.xyz2, sm.off, q0, p.qt, total_q, tmp, ctr, nullptr, nullptr);
        compute_sync();
        for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
            int ql, k;
            rows.decode(r, ql, k);
            nbr[r] = (ql >= 0 && ql < p.qt) ? tmp[ql * g.K + k] : -1;
        }
        compute_sync();
    }
    const int m = pipe.my_row();
    if (p.nbr_in[set] != nullptr && g_direct_gather_dev(p)) {
        // Gather straight into tensor memory: a compute thread owns row m of the tile and every other 16-column block
        // of the A operand, so it can fetch exactly the channels it will write -- the neighbour's xyz and 16-float
        // pieces of its feature row -- in ONE round trip (all loads issued before any is used), split them and store
        // them with tcgen05.st.  No staging in shared memory, no transposition, no second trip for the features.
        int ql, kk;
        rows.decode(m, ql, kk);
        int bq = -1;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        const float4* frow = nullptr;
        if (ql >= 0 && ql < p.qt) bq = __float_as_int(ctr[ql * 4 + 3]);
        if (bq >= 0) {
            const int cell = nbr[m];
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (cell >= 0) {
                const float* sx = p.xyz2 + ((size_t)bq * cells2 + cell) * 3;
                qx = __ldg(sx); qy = __ldg(sx + 1); qz = __ldg(sx + 2);
                frow = reinterpret_cast<const float4*>(p.feat2[set] + ((size_t)bq * cells2 + cell) * p.Cf);
            }
            d0 = qx - ctr[ql * 4 + 0]; d1 = qy - ctr[ql * 4 + 1]; d2 = qz - ctr[ql * 4 + 2];
        }
        const int nblk = (cin0 + 15) >> 4, ng = p.Cf >> 2;
        constexpr int TB = 3;                    // blocks per trip: 15 float4 in flight (Cf = 64: one trip)
        bool allocated = false;
        for (int b0 = pipe.my_half(); b0 < nblk || !allocated; b0 += 2 * TB) {
            // channel c = 16 b + i is feature c - 3: the five float4 groups 4b-1 .. 4b+3 cover a block, shifted by one
            float4 L[TB][5];
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const int b = b0 + 2 * u;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int gi = 4 * b - 1 + j;
                    L[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b < nblk && frow != nullptr && gi >= 0 && gi < ng) L[u][j] = __ldg(frow + gi);
                }
            }
            if (!allocated) { pipe.alloc_late(); allocated = true; }   // tensor memory is claimed with the loads in flight
            const uint32_t addr = pipe.my_lane_addr();                 // (its base address is known only now)
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const int b = b0 + 2 * u;
                if (b >= nblk) continue;
                const float* flat = reinterpret_cast<const float*>(&L[u][0]);
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v = flat[i + 1];
                    if (b == 0 && i < 3) v = i == 0 ? d0 : (i == 1 ? d1 : d2);
                    tc::split_tf32(v, hi[i], lo[i]);
                }
                tc::tmem_st16(addr + b * 16, hi);
                tc::tmem_st16(addr + TC_A_LO + b * 16, lo);
            }
        }
    } else {
    for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
        float pc[3], qc[3];
        tc_row_xyz(rows, r, p.qt, nbr, ctr, p.xyz2, cells2, pc, qc);
        X[act_index(0, r, RS)] = qc[0] - pc[0];
        X[act_index(1, r, RS)] = qc[1] - pc[1];
        X[act_index(2, r, RS)] = qc[2] - pc[2];
    }
    gather_features(X, RS, 3, p.feat2[set], p.Cf, RS, [&](int r) -> long long {
        int ql, k;
        rows.decode(r, ql, k);
        if (ql < 0 || ql >= p.qt || nbr[r] < 0) return -1;
        return (long long)__float_as_int(ctr[ql * 4 + 3]) * cells2 + nbr[r];
    });
    pipe.alloc_late();                           // (a barrier of the compute warps: the gather above is complete)
    pipe.load_a_from_smem(X, 0, cin0, 0);
    }
    pipe.signal_a_ready();
    float* S = X;                                // staging reused as S[row][cout_last + 1]: X is dead once loaded
    for (int l = 0; l < p.nl; ++l) {
        const bool last = l == p.nl - 1;
        pipe.epilogue<true>(p.cout[l], [&](int b, const float (&v)[16]) {
            if (!last) { pipe.store_a(0, b, v); return; }
#pragma unroll
            for (int i = 0; i < 16; ++i) S[m * pool_ld + b * 16 + i] = v[i];
        });
        if (!last) pipe.signal_a_ready();
    }
    pipe.finish();                               // (a barrier too) tensor memory goes to the SM's other tile now
    // max over the K neighbours of (y * mask): y >= 0 after ReLU, masked rows count as 0
    float* out = p.out[set];
    for (int t = threadIdx.x; t < p.qt * cout_last; t += CTA_THREADS) {
        const int ql = t / cout_last, c = t - ql * cout_last;
        const long long gq = q0 + ql;
        if (gq >= total_q) break;
        const int r0 = rows.row(ql, 0);
        float mx = 0.f;
        for (int k = 0; k < g.K; ++k)
            if (nbr[r0 + k] >= 0) mx = fmaxf(mx, S[(r0 + k) * pool_ld + c]);
        out[gq * cout_last + c] = mx;
    }
    if (p.dbg_nbr[set] != nullptr)
        for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
            int q2, k2;
            rows.decode(r, q2, k2);
            if (q2 >= 0 && q2 < p.qt && q0 + q2 < total_q) p.dbg_nbr[set][(q0 + q2) * g.K + k2] = nbr[r];
        }
}

__global__ void __launch_bounds__(TC_LAUNCH_THREADS, 2) cost_volume_1_tc_kernel(const Cv1Params p)
{
    constexpr int RS = TC_ROWS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C, xc = 10 + 2 * C;
    TcSmem sm = tc_carve(smem_raw, p.nring, 0, (size_t)max(((xc + 3) & ~3), (p.sl_in_ring ? 1 : 2) * POOL_LD64) * RS);
    const int warp = threadIdx.x >> 5;
    const Window g = p.g;
    const TcRows rows(g.K);
    int* nbr = sm.nbr; float* ctr = sm.ctr; float* X = sm.X;
    const long long q0 = p.q_first + (long long)blockIdx.x * p.qt;
    const int cells = g.h2 * g.w2;
    pdl_trigger();
    TcPipe pipe;
    pipe.begin(sm.ring, sm.bars, p.nring, sm.tmem_holder, p.weights + (size_t)p.total_chunks * TC_CHUNK_FLOATS,
               sm.bias, 128 + 64 + 64 + 64 + 128 + 64, p.tlog, /*alloc_now=*/false);
    if (warp < COMPUTE_WARPS) {
        pdl_wait();
        tc_load_nbr(p.qs, rows, p.xyz1, p.nbr_in, q0, p.qt, p.total_q, nbr, ctr);
    }
    pipe.join();
    if (warp == COMPUTE_WARPS) { pipe.produce(p.weights, p.total_chunks); return; }
    if (warp == COMPUTE_WARPS + 1) {
        if (tc::elect_one()) {
            pipe.issue_layer(0, xc, 128);     // CV_0
            pipe.issue_layer(0, 128, 64);     // CV_1
            pipe.issue_layer(0, 64, 64);      // CV_2      -> F
            pipe.issue_layer(0, 10, 64);      // CV_xyz    -> enc
            pipe.issue_layer(0, 128, 128);    // sum_CV_0 on [enc | F]
            pipe.issue_layer(0, 128, 64);     // sum_CV_1  -> logits
        }
        return;
    }
    pipe.stamp(2);
    const int m = pipe.my_row();
    // the 10 xyz channels are needed again by CV_xyz after the tile's A region has been overwritten
    float x10[16];
    if (p.direct_gather) {
        // Gather straight into tensor memory (see group_mlp_max_tc_kernel): row m's [p, q, q - p, |q - p|] from the
        // centre table and one xyz load, and 16-float pieces of the virtual vector [f1 of the query | f2 of the
        // neighbour] -- channel c >= 10 is element c - 10 of it, so a 16-channel block is covered by five float4 groups
        // shifted by two -- all in ONE round trip, then tcgen05.st.
        int ql, kk;
        rows.decode(m, ql, kk);
        const int nq = p.qs.oh * p.qs.ow, ng = C >> 2;
        const bool okq = ql >= 0 && ql < p.qt && q0 + ql < p.total_q;
        int bq = -1;
        if (ql >= 0 && ql < p.qt) bq = __float_as_int(ctr[ql * 4 + 3]);
        float px = 0.f, py = 0.f, pz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
        const int cell = (ql >= 0 && ql < p.qt) ? nbr[m] : -1;
        if (bq >= 0) {
            px = ctr[ql * 4 + 0]; py = ctr[ql * 4 + 1]; pz = ctr[ql * 4 + 2];
            if (cell >= 0) {
                const float* sx = p.xyz2 + ((size_t)bq * cells + cell) * 3;
                qx = __ldg(sx); qy = __ldg(sx + 1); qz = __ldg(sx + 2);
            }
        }
        const float4* f1row = okq ? reinterpret_cast<const float4*>(p.f1 + (size_t)(q0 + ql) * C) : nullptr;
        const float4* f2row = (ql >= 0 && ql < p.qt && cell >= 0)
            ? reinterpret_cast<const float4*>(p.f2 + ((size_t)((q0 + ql) / nq) * cells + cell) * C) : nullptr;
        const int nblk = (xc + 15) >> 4;
        constexpr int TB = 3;
        bool allocated = false;
        for (int b0 = pipe.my_half(); b0 < nblk || !allocated; b0 += 2 * TB) {
            float4 L[TB][5];
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const int b = b0 + 2 * u;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int gi = 4 * b - 3 + j;
                    L[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b < nblk && gi >= 0) {
                        if (gi < ng) { if (f1row != nullptr) L[u][j] = __ldg(f1row + gi); }
                        else if (gi < 2 * ng) { if (f2row != nullptr) L[u][j] = __ldg(f2row + (gi - ng)); }
                    }
                }
            }
            if (!allocated) {
                const float dx = qx - px, dy = qy - py, dz = qz - pz;
                x10[0] = px; x10[1] = py; x10[2] = pz; x10[3] = qx; x10[4] = qy; x10[5] = qz;
                x10[6] = dx; x10[7] = dy; x10[8] = dz;
                x10[9] = sqrtf(__fadd_rn(sumsq_tf(dx, dy, dz), 1e-20f));
#pragma unroll
                for (int i = 10; i < 16; ++i) x10[i] = 0.f;
                pipe.alloc_late();               // tensor memory is claimed with the loads in flight
                pipe.stamp(3);
                allocated = true;
            }
            const uint32_t addr = pipe.my_lane_addr();
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const int b = b0 + 2 * u;
                if (b >= nblk) continue;
                const float* flat = reinterpret_cast<const float*>(&L[u][0]);
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v = flat[i + 2];
                    if (b == 0 && i < 10) v = x10[i];
                    tc::split_tf32(v, hi[i], lo[i]);
                }
                tc::tmem_st16(addr + b * 16, hi);
                tc::tmem_st16(addr + TC_A_LO + b * 16, lo);
            }
        }
    } else {
    for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
        float pc[3], qc[3];
        tc_row_xyz(rows, r, p.qt, nbr, ctr, p.xyz2, cells, pc, qc);
        write_xyz10(X, RS, r, pc[0], pc[1], pc[2], qc[0], qc[1], qc[2]);
    }
    {
        const int qt = p.qt, nq = p.qs.oh * p.qs.ow;
        const long long total = p.total_q;
        // f1 of the query pixel, repeated for its K rows (queries are all pixels: linear cell = gq)
        gather_features(X, RS, 10, p.f1, C, RS, [&](int r) -> long long {
            int ql, k;
            rows.decode(r, ql, k);
            return (ql >= 0 && ql < qt && q0 + ql < total) ? q0 + ql : -1;
        });
        gather_features(X, RS, 10 + C, p.f2, C, RS, [&](int r) -> long long {
            int ql, k;
            rows.decode(r, ql, k);
            if (ql < 0 || ql >= qt || nbr[r] < 0) return -1;
            return (long long)((q0 + ql) / nq) * cells + nbr[r];
        });
    }
    pipe.alloc_late();                          // (a barrier of the compute warps: the gather above is complete)
    pipe.stamp(3);
#pragma unroll
    for (int i = 0; i < 16; ++i) x10[i] = i < 10 ? X[act_index(i, m, RS)] : 0.f;
    pipe.load_a_from_smem(X, 0, xc, 0);
    }
    pipe.signal_a_ready();
    pipe.stamp(4);
    // pool staging (the X region is dead once every thread has loaded its row): F and logits as [row][65]
    // The logits are written by the LAST epilogue, when every MMA has completed and the weight ring is dead:
    // with sl_in_ring they go there, which keeps the tile small enough for two to be resident per SM.
    float* SF = X;
    float* SL = p.sl_in_ring ? sm.ring : X + RS * POOL_LD64;
    pipe.epilogue<true>(128, [&](int b, const float (&v)[16]) { pipe.store_a(0, b, v); });        // CV_0
    pipe.signal_a_ready();
    pipe.epilogue<true>(64, [&](int b, const float (&v)[16]) { pipe.store_a(0, b, v); });         // CV_1
    pipe.signal_a_ready();
    pipe.epilogue<true>(64, [&](int b, const float (&v)[16]) {                                    // CV_2 = F
        pipe.store_a(64, b, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) SF[m * POOL_LD64 + b * 16 + i] = v[i];
    });
    if (pipe.my_half() == 0) {                  // re-materialise the xyz block for CV_xyz
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) tc::split_tf32(x10[i], hi[i], lo[i]);
        tc::tmem_st16(pipe.my_lane_addr(), hi);
        tc::tmem_st16(pipe.my_lane_addr() + TC_A_LO, lo);
    }
    pipe.signal_a_ready();
    pipe.epilogue<true>(64, [&](int b, const float (&v)[16]) { pipe.store_a(0, b, v); });         // CV_xyz = enc
    pipe.signal_a_ready();
    pipe.epilogue<true>(128, [&](int b, const float (&v)[16]) { pipe.store_a(0, b, v); });        // sum_CV_0
    pipe.signal_a_ready();
    pipe.epilogue<true>(64, [&](int b, const float (&v)[16]) {                                    // logits
#pragma unroll
        for (int i = 0; i < 16; ++i) SL[m * POOL_LD64 + b * 16 + i] = v[i];
    });
    pipe.finish();                              // (a barrier too) tensor memory goes to the SM's other tile now
    pipe.stamp(5);
    for (int t = threadIdx.x; t < p.qt * 64; t += CTA_THREADS) {
        const int ql = t >> 6, c = t & 63;
        const long long gq = q0 + ql;
        if (gq >= p.total_q) break;
        p.out[gq * 64 + c] = pool_softmax(SL, SF, POOL_LD64, nbr, rows.row(ql, 0), g.K, c);
    }
    pipe.stamp(6);
    if (p.dbg_nbr != nullptr)
        for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
            int q2, k2;
            rows.decode(r, q2, k2);
            if (q2 >= 0 && q2 < p.qt && q0 + q2 < p.total_q) p.dbg_nbr[(q0 + q2) * g.K + k2] = nbr[r];
        }
    pipe.stamp(7);
}

__global__ void __launch_bounds__(TC_LAUNCH_THREADS, 2) cost_volume_2_tc_kernel(const Cv2Params p)
{
    constexpr int RS = TC_ROWS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C;
    // staging channels: [0,10) xyz, [16,16+C) f1, [16+C,80+C) stage-1 features of the neighbour; then the pool
    // staging SV / SL as [row][65]
    TcSmem sm = tc_carve(smem_raw, p.nring, 0, (size_t)(80 + C + (p.sl_in_ring ? 0 : 2 * POOL_LD64)) * RS);
    const int warp = threadIdx.x >> 5;
    const Window g = p.g;
    const TcRows rows(g.K);
    int* nbr = sm.nbr; float* ctr = sm.ctr; float* X = sm.X;
    const long long q0 = p.q_first + (long long)blockIdx.x * p.qt;
    const int cells = g.h2 * g.w2;
    pdl_trigger();
    TcPipe pipe;
    pipe.begin(sm.ring, sm.bars, p.nring, sm.tmem_holder, p.weights + (size_t)p.total_chunks * TC_CHUNK_FLOATS,
               sm.bias, 64 + 128 + 64, p.tlog, /*alloc_now=*/false);
    if (warp < COMPUTE_WARPS) {
        pdl_wait();
        tc_load_nbr(p.qs, rows, p.xyz1, p.nbr_in, q0, p.qt, p.total_q, nbr, ctr);
    }
    pipe.join();
    if (warp == COMPUTE_WARPS) { pipe.produce(p.weights, p.total_chunks); return; }
    if (warp == COMPUTE_WARPS + 1) {
        if (tc::elect_one()) {
            pipe.issue_layer(0, 10, 64);          // sum_xyz_encoding
            pipe.issue_layer(0, 128 + C, 128);    // sum_cost_volume_0 on [enc | f1 | stage-1]
            pipe.issue_layer(0, 128, 64);         // sum_cost_volume_1 -> logits
        }
        return;
    }
    for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
        float pc[3], qc[3];
        tc_row_xyz(rows, r, p.qt, nbr, ctr, p.xyz1, cells, pc, qc);
        write_xyz10(X, RS, r, pc[0], pc[1], pc[2], qc[0], qc[1], qc[2]);
    }
    {
        const int qt = p.qt, nq = p.qs.oh * p.qs.ow;
        const long long total = p.total_q;
        gather_features(X, RS, 16, p.f1, C, RS, [&](int r) -> long long {
            int ql, k;
            rows.decode(r, ql, k);
            return (ql >= 0 && ql < qt && q0 + ql < total) ? q0 + ql : -1;
        });
        gather_features<8>(X, RS, 16 + C, p.cv1, 64, RS, [&](int r) -> long long {
            int ql, k;
            rows.decode(r, ql, k);
            if (ql < 0 || ql >= qt || nbr[r] < 0) return -1;
            return (long long)((q0 + ql) / nq) * cells + nbr[r];
        });
    }
    pipe.alloc_late();                          // (a barrier of the compute warps: the gather above is complete)
    pipe.load_a_from_smem(X, 0, 10, 0);
    pipe.load_a_from_smem(X, 16, C + 64, 64);
    pipe.signal_a_ready();
    const int m = pipe.my_row();
    // Pool staging.  sl_in_ring (two tiles per SM): the logits go to the weight ring, dead by the time the last
    // epilogue writes them, and the values are pooled straight from their place in X (4-way bank conflicts on
    // K loads per output); otherwise both get a [row][65] copy behind X.
    float* SV = X + (size_t)(80 + C) * RS;
    float* SL = p.sl_in_ring ? sm.ring : SV + RS * POOL_LD64;
    if (!p.sl_in_ring)      // while the first MMAs run: this row's gathered stage-1 features (masked rows hold 0)
        for (int c = pipe.my_half() * 32; c < pipe.my_half() * 32 + 32; ++c) SV[m * POOL_LD64 + c] = X[act_index(16 + C + c, m, RS)];
    pipe.epilogue<true>(64, [&](int b, const float (&v)[16]) { pipe.store_a(0, b, v); });         // enc
    pipe.signal_a_ready();
    pipe.epilogue<true>(128, [&](int b, const float (&v)[16]) { pipe.store_a(0, b, v); });
    pipe.signal_a_ready();
    pipe.epilogue<true>(64, [&](int b, const float (&v)[16]) {                                    // logits
#pragma unroll
        for (int i = 0; i < 16; ++i) SL[m * POOL_LD64 + b * 16 + i] = v[i];
    });
    pipe.finish();                              // (a barrier too) tensor memory is free while this tile pools
    for (int t = threadIdx.x; t < p.qt * 64; t += CTA_THREADS) {
        const int ql = t >> 6, c = t & 63;
        const long long gq = q0 + ql;
        if (gq >= p.total_q) break;
        p.out[gq * 64 + c] = p.sl_in_ring
            ? pool_softmax_v(SL, [&](int row, int ch) { return X[act_index(16 + C + ch, row, RS)]; }, nbr, rows.row(ql, 0), g.K, c)
            : pool_softmax(SL, SV, POOL_LD64, nbr, rows.row(ql, 0), g.K, c);
    }
    if (p.dbg_nbr != nullptr)
        for (int r = threadIdx.x; r < RS; r += CTA_THREADS) {
            int q2, k2;
            rows.decode(r, q2, k2);
            if (q2 >= 0 && q2 < p.qt && q0 + q2 < p.total_q) p.dbg_nbr[(q0 + q2) * g.K + k2] = nbr[r];
        }
}

__global__ void __launch_bounds__(TC_LAUNCH_THREADS, 2) row_mlp_tc_kernel(const RowMlpParams p)
{
    constexpr int RS = TC_ROWS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int set = blockIdx.y, warp = threadIdx.x >> 5;
    TcSmem sm = tc_carve(smem_raw, p.nring, 0, (size_t)p.stage_ch * RS);
    pdl_trigger();
    TcPipe pipe;
    int nbias = 0;
    for (int ph = 0; ph < p.nphase; ++ph)
        for (int l = 0; l < p.nl[ph]; ++l) nbias += p.cout[ph][l];
    // Tensor memory is claimed only after the dependent-launch wait and the gather (alloc_late): no tile of this
    // library ever holds it while waiting for another kernel, whatever shares the SM with it.
    pipe.begin(sm.ring, sm.bars, p.nring, sm.tmem_holder, p.weights[set] + (size_t)p.total_chunks * TC_CHUNK_FLOATS,
               sm.bias, nbias, p.tlog, /*alloc_now=*/false);
    pipe.join();
    if (warp == COMPUTE_WARPS) { pipe.produce(p.weights[set], p.total_chunks); return; }
    if (warp == COMPUTE_WARPS + 1) {
        if (tc::elect_one())
            for (int ph = 0; ph < p.nphase; ++ph) {
                int cin = 0;
                for (int i = 0; i < p.nsrc[ph]; ++i) cin += p.src_c[ph][i];
                for (int l = 0; l < p.nl[ph]; ++l) { pipe.issue_layer(0, cin, p.cout[ph][l]); cin = p.cout[ph][l]; }
            }
        return;
    }
    pdl_wait();
    float* X = sm.X;
    const long long r0 = (long long)blockIdx.x * p.rt;
    const long long rows = p.rows;
    const int rt = p.rt;
    for (int ph = 0; ph < p.nphase; ++ph)
        for (int i = 0; i < p.nsrc[ph]; ++i) {
            if (p.src_prev[ph][i] || !p.stage_here[ph][i]) continue;    // a source used twice is staged once
            {
                // 64 channels x 128 rows are 8 float4 per thread: one round trip per source instead of two
                gather_features<8>(X, RS, p.slot[ph][i], p.src[set][ph][i], p.src_c[ph][i], RS, [&](int r) -> long long {
                    return (r < rt && r0 + r < rows) ? r0 + r : -1;
                });
            }
        }
    pipe.alloc_late();                          // (a barrier of the compute warps: the gather above is complete)
    const int m = pipe.my_row();
    const bool row_ok = m < rt && r0 + m < rows;
    for (int ph = 0; ph < p.nphase; ++ph) {
        // global sources of this phase -> their place in the concatenation
        int off = 0, prev_off_next = 0;
        for (int i = 0; i < p.nsrc[ph]; ++i) {
            if (!p.src_prev[ph][i]) pipe.load_a_from_smem(X, p.slot[ph][i], p.src_c[ph][i], off);
            off += p.src_c[ph][i];
        }
        if (ph + 1 < p.nphase) {
            int o = 0;
            for (int i = 0; i < p.nsrc[ph + 1]; ++i) { if (p.src_prev[ph + 1][i]) prev_off_next = o; o += p.src_c[ph + 1][i]; }
        }
        pipe.signal_a_ready();
        for (int l = 0; l < p.nl[ph]; ++l) {
            const bool last_layer = l == p.nl[ph] - 1;
            const bool last_phase = ph == p.nphase - 1;
            const int n = p.cout[ph][l];
            float* gout = !last_layer ? nullptr : (last_phase ? p.out[set] : p.out_phase0[set]);
            pipe.epilogue<true>(n, [&](int b, const float (&v)[16]) {
                if (!last_layer) pipe.store_a(0, b, v);
                else if (!last_phase) pipe.store_a(prev_off_next, b, v);
                if (gout != nullptr && row_ok) {
                    float4* dst = reinterpret_cast<float4*>(gout + (r0 + m) * n + b * 16);
                    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                    dst[2] = make_float4(v[8], v[9], v[10], v[11]);
                    dst[3] = make_float4(v[12], v[13], v[14], v[15]);
                }
            });
            if (!last_layer) pipe.signal_a_ready();
            // after the last layer of a non-final phase the next iteration loads the other sources and signals
        }
    }
    pipe.finish();
}

// ================================================================================================
// host side
struct TileChoice { int nb, per_tile, tiles; };

// Pick rows-per-tile (64 or 128) and a balanced split: `units` work items of `rows_per_unit` rows each.
template <typename SmemFn>
static TileChoice choose_tile(long long units, int rows_per_unit, int nsets, SmemFn smem_bytes)
{
    const int sms = device_info().sm_count;
    TileChoice best{0, 0, 0};
    double best_cost = 1e30;
    for (int nb = 1; nb <= 2; ++nb) {
        if (smem_bytes(nb) + 2 * CHUNK_BYTES > (size_t)SMEM_LIMIT) continue;
        const int cap = (64 * nb) / rows_per_unit;
        if (cap < 1) continue;
        long long tiles = (units + cap - 1) / cap;
        const long long waves = (tiles * nsets + sms - 1) / sms;
        // one 128-row tile costs ~1.7x a 64-row tile (weights and barriers are amortised over more rows)
        const double cost = (double)waves * (nb == 1 ? 1.0 : 1.7);
        if (cost < best_cost) {
            best_cost = cost;
            // spread the units evenly over as many tiles as the chosen number of waves can hold
            long long slots = waves * sms / nsets;
            if (slots < tiles) slots = tiles;
            if (slots > units) slots = units;
            int per = (int)((units + slots - 1) / slots);
            if (per > cap) per = cap;
            best = TileChoice{nb, per, (int)((units + per - 1) / per)};
        }
    }
    return best;
}

template <typename Kernel>
static int set_smem(Kernel k, size_t bytes, const char* what)
{
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err != cudaSuccess) return set_cuda_error(err, what);
    // all of the SM's L1/shared array as shared memory: two ~110 KB tiles are meant to be resident together
    err = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return set_cuda_error(err, what);
    return 0;
}

static size_t align16(size_t b) { return (b + 15) & ~size_t(15); }

// shared memory besides the weight ring: barriers, offset table, row->cell table, centres
static size_t common_smem(int kt, int rs)
{
    return align16(2 * MAX_RING * 8) + align16((size_t)kt * 8) + align16((size_t)rs * 4) + align16((size_t)rs * 16);
}

// ring depth: as many 8 KB chunks as fit next to `base` bytes, 2..MAX_RING, no more than the stream has
static int pick_ring(size_t base, int total_chunks)
{
    long long n = ((long long)SMEM_LIMIT - (long long)base) / CHUNK_BYTES;
    if (n > MAX_RING) n = MAX_RING;
    if (n > total_chunks) n = total_chunks;
    return (int)n;
}

static int check_window(const elo_window* w, const char* who)
{
    if (w->kernel_size_H <= 0 || w->kernel_size_W <= 0 || w->K <= 0 || !(w->distance > 0) || w->stride_h <= 0 ||
        w->stride_w <= 0 || w->small_h <= 0 || w->small_w <= 0 || w->random_hw == nullptr)
        return set_error(ELO_ERR_INVALID_ARGUMENT, who);
    if ((long long)w->kernel_size_H * w->kernel_size_W > 5000)
        return set_error(ELO_ERR_UNSUPPORTED, "window larger than 5000 cells");
    if (w->small_h >= 32768 || w->small_w >= 32768) return set_error(ELO_ERR_UNSUPPORTED, "grid extent > 32767");
    return 0;
}

static Window make_window(const elo_window* w)
{
    Window g;
    g.h2 = w->small_h; g.w2 = w->small_w; g.kH = w->kernel_size_H; g.kW = w->kernel_size_W;
    g.kt = w->kernel_size_H * w->kernel_size_W; g.stride_h = w->stride_h; g.stride_w = w->stride_w;
    g.K = w->K; g.flag_copy = 0; g.d2max = w->distance * w->distance;
    return g;
}

static QuerySet make_queries(const elo_queries* q)
{
    QuerySet s;
    s.H1 = q->H; s.W1 = q->W; s.oh = q->out_h; s.ow = q->out_w; s.qs_h = q->q_stride_h; s.qs_w = q->q_stride_w;
    return s;
}

static int check_queries(const elo_queries* q, const char* who)
{
    if (q->H <= 0 || q->W <= 0 || q->out_h <= 0 || q->out_w <= 0 || q->q_stride_h <= 0 || q->q_stride_w <= 0 ||
        (q->out_h - 1) * q->q_stride_h >= q->H || (q->out_w - 1) * q->q_stride_w >= q->W)
        return set_error(ELO_ERR_INVALID_ARGUMENT, who);
    return 0;
}

static int width_ok(int c) { return c == 64 || c == 128; }

static long long* g_tlog = nullptr;    // optional device buffer of 64 timestamps (elo_set_time_log)

// MLP engine: 1 = tcgen05 tensor cores (3xTF32, 128-row tiles), 0 = fp32 FFMA (64/128-row tiles)
static int g_engine = 1;

struct TcChoice { int per_tile, tiles; };

static int g_tile_policy = (getenv("ELO_TILE_POLICY") && getenv("ELO_TILE_POLICY")[0] == '1') ? 1 : 0;   // 0: latency (spread over the SMs), 1: throughput (full tiles)

// 128-row tiles: `units` work items of `rows_per_unit` rows, spread evenly over the waves they need
static TcChoice choose_tc_tile(long long units, int rows_per_unit, int nsets, int ctas_per_sm = 1)
{
    // a query's rows never straddle a 32-row lane quarter (TcRows): 4 * floor(32 / K) queries per tile
    const int cap = rows_per_unit > 1 ? 4 * (32 / rows_per_unit) : TC_ROWS;
    long long tiles = (units + cap - 1) / cap;
    // Two tiles resident on an SM take turns on its tensor memory: that pays when the work needs more than one
    // wave of full tiles anyway; a call that fits one tile per SM keeps the SMs to itself.
    // Throughput policy (elo_set_tile_policy(1)): full tiles, as few CTAs as the work needs.  A tile's latency
    // hardly depends on how many of its 128 rows are in use, so spreading a small call over all SMs (the
    // default, latency policy) buys a little latency with a lot of SM time -- time that other forwards in
    // flight on other streams could use.
    if (g_tile_policy == 1) return TcChoice{cap, (int)tiles};
    int sms = device_info().sm_count;
    if (tiles * nsets > sms) sms *= ctas_per_sm;
    const long long waves = (tiles * nsets + sms - 1) / sms;
    long long slots = waves * sms / nsets;
    if (slots < tiles) slots = tiles;
    if (slots > units) slots = units;
    int per = (int)((units + slots - 1) / slots);
    if (per > cap) per = cap;
    return TcChoice{per, (int)((units + per - 1) / per)};
}

// shared memory of a tensor-core tile besides the weight ring
static size_t tc_base_smem(int kt, size_t staging_floats)
{
    return align16((2 * MAX_RING + 2) * 8) + 16 + align16(TC_MAX_BIAS * 4) + align16((size_t)kt * 8) +
           align16(TC_ROWS * 4) + align16(TC_ROWS * 16) + align16(staging_floats * 4);
}

// shared memory one CTA may use when two are to be resident on an SM (228 KB per SM, 1 KB reserved per CTA,
// 1 KB of static shared memory in these kernels)
static constexpr int SMEM_HALF = 112 * 1024;

static int g_ring_alone = getenv("ELO_TC_RING") ? atoi(getenv("ELO_TC_RING")) : 3;     // measured: 3..8 slots give the same tile latency

static int g_ring_shared = getenv("ELO_TC_RING_SHARED") ? atoi(getenv("ELO_TC_RING_SHARED")) : 3;

static int tc_pick_ring(size_t base, int total_chunks, bool two_per_sm = false)
{
    if (two_per_sm) {
        long long n2 = ((long long)SMEM_HALF - (long long)base) / TC_CHUNK_BYTES;
        if (n2 > total_chunks) n2 = total_chunks;
        if (n2 > g_ring_shared && g_ring_shared >= 2) n2 = g_ring_shared;
        if (n2 >= 2 || n2 == total_chunks) return (int)(n2 > MAX_RING ? MAX_RING : n2);
    }
    long long n = ((long long)SMEM_LIMIT - (long long)base) / TC_CHUNK_BYTES;
    // a tile that has the SM to itself still leaves room for the small kernels of other forwards in flight
    // (searches, projections, pose heads): TC_RING_ALONE slots keep the MMA warp fed (g_ring_alone, tunable)
    if (n > g_ring_alone) n = g_ring_alone;
    if (n > MAX_RING) n = MAX_RING;
    if (n > total_chunks) n = total_chunks;
    return (int)n;
}

template <typename Kernel, typename Params>
static int launch_tc(Kernel kern, const Params& p, dim3 grid, size_t bytes, cudaStream_t st, const char* what)
{
    int rc = set_smem(kern, bytes, what);
    if (rc) return rc;
    cudaError_t err = launch(kern, grid, dim3(TC_LAUNCH_THREADS), bytes, st, p);
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, what);
}

}  // namespace elo

using namespace elo;

extern "C" int elo_set_tile_policy(int policy)
{
    if (policy != 0 && policy != 1) return set_error(ELO_ERR_INVALID_ARGUMENT, "elo_set_tile_policy: 0 (latency) or 1 (throughput)");
    g_tile_policy = policy;
    return ELO_OK;
}

extern "C" int elo_get_tile_policy(void) { return g_tile_policy; }

extern "C" int elo_set_time_log(long long* device_buf)
{
    g_tlog = device_buf;
    return ELO_OK;
}

extern "C" int elo_set_mlp_engine(int engine)
{
    if (engine != 0 && engine != 1) return set_error(ELO_ERR_INVALID_ARGUMENT, "mlp engine: 0 = fp32 FFMA, 1 = tcgen05 3xTF32");
    g_engine = engine;
    return ELO_OK;
}
extern "C" int elo_get_mlp_engine(void) { return g_engine; }

extern "C" int elo_group_mlp_max(const elo_group_mlp_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "group_mlp_max: null descriptor");
    int rc = check_window(&d->window[0], "group_mlp_max: bad window");
    if (rc) return rc;
    rc = check_queries(&d->queries, "group_mlp_max: bad query grid");
    if (rc) return rc;
    if (d->batch_size < 0 || d->nsets < 1 || d->nsets > 2 || d->num_layers < 1 || d->num_layers > 3 ||
        d->feat_channels <= 0 || (d->feat_channels & 3) || !d->xyz1 || !d->xyz2)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "group_mlp_max: bad arguments");
    for (int l = 0; l < d->num_layers; ++l)
        if (!width_ok(d->cout[l])) return set_error(ELO_ERR_UNSUPPORTED, "group_mlp_max: layer widths must be 64 or 128");
    if (d->window[0].K > 64) return set_error(ELO_ERR_UNSUPPORTED, "group_mlp_max: K > 64");
    if (d->nsets == 2) {
        // the two sets of a launch share ONE window geometry (only the scan order, features and weights differ)
        const elo_window &a = d->window[0], &b = d->window[1];
        if (a.kernel_size_H != b.kernel_size_H || a.kernel_size_W != b.kernel_size_W || a.K != b.K ||
            a.distance != b.distance || a.stride_h != b.stride_h || a.stride_w != b.stride_w ||
            a.small_h != b.small_h || a.small_w != b.small_w)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "group_mlp_max: window[1] differs from window[0] in geometry");
    }
    if (d->batch_size == 0) return ELO_OK;

    GroupMlpParams p;
    static const int direct_gather = getenv("ELO_DIRECT_GATHER") ? atoi(getenv("ELO_DIRECT_GATHER")) : 1;
    p.direct_gather = direct_gather;
    p.qs = make_queries(&d->queries);
    p.g = make_window(&d->window[0]);
    long long per_set = (long long)d->batch_size * p.qs.oh * p.qs.ow;
    long long qb = 0;                        // optional sub-range of each set's queries (row bands)
    if (d->query_begin != 0 || d->query_end != 0) {
        if (d->query_begin < 0 || d->query_end > per_set || d->query_begin > d->query_end)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "group_mlp_max: bad query range");
        qb = d->query_begin;
        per_set = d->query_end - d->query_begin;
        if (per_set == 0) return ELO_OK;
    }
    p.Cf = d->feat_channels;
    p.nl = d->num_layers;
    int cin = 3 + p.Cf, chunks = 0;
    for (int l = 0; l < 3; ++l) {
        p.cout[l] = l < p.nl ? d->cout[l] : 0;
        if (l < p.nl) { chunks += layer_chunks(cin, p.cout[l]); cin = p.cout[l]; }
    }
    p.total_chunks = chunks;
    p.xyz1 = d->xyz1; p.xyz2 = d->xyz2;
    for (int s = 0; s < 2; ++s) {
        const int u = s < d->nsets ? s : 0;
        if (!d->feat2[u] || !d->weights[u] || !d->out[u] || !d->window[u].random_hw)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "group_mlp_max: null pointer");
        p.feat2[s] = d->feat2[u]; p.random_hw[s] = d->window[u].random_hw; p.weights[s] = d->weights[u];
        p.out[s] = d->out[u]; p.dbg_nbr[s] = d->dbg_nbr[u]; p.nbr_in[s] = d->nbr[u]; p.tlog = g_tlog;
        if (d->set_batch_offset[u] < 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "group_mlp_max: negative batch offset");
        p.q_base[s] = (long long)d->set_batch_offset[u] * p.qs.oh * p.qs.ow + qb;
        p.q_end[s] = p.q_base[s] + per_set;
    }
    const int kt = p.g.kt, xch = (3 + p.Cf + 3) & ~3;
    if (g_engine == 1) {
        if (p.g.K > 32) return set_error(ELO_ERR_UNSUPPORTED, "group_mlp_max: the tensor-core engine takes K <= 32");
        int cin_t = 3 + p.Cf, chunks_t = 0;
        for (int l = 0; l < p.nl; ++l) { chunks_t += tc_layer_chunks(cin_t, p.cout[l]); cin_t = p.cout[l]; }
        p.total_chunks = chunks_t;
        const int pool_ld = p.cout[p.nl - 1] + 1;
        const size_t base = tc_base_smem(kt, (size_t)(xch > pool_ld ? xch : pool_ld) * TC_ROWS);
        p.nring = tc_pick_ring(base, chunks_t, /*two_per_sm=*/true);
        if (p.nring < 2) return set_error(ELO_ERR_UNSUPPORTED, "group_mlp_max: tile does not fit shared memory");
        const bool two = base + (size_t)p.nring * TC_CHUNK_BYTES <= (size_t)SMEM_HALF;
        const TcChoice tc = choose_tc_tile(per_set, p.g.K, d->nsets, two ? 2 : 1);
        p.qt = tc.per_tile;
        return launch_tc(group_mlp_max_tc_kernel, p, dim3(tc.tiles, d->nsets), base + (size_t)p.nring * TC_CHUNK_BYTES,
                         (cudaStream_t)stream, "group_mlp_max (tensor core) launch");
    }
    auto smem = [&](int nb) { return common_smem(kt, 64 * nb) + (size_t)(xch + 256) * 64 * nb * 4; };
    const TileChoice tc = choose_tile(per_set, p.g.K, d->nsets, smem);
    if (tc.nb == 0) return set_error(ELO_ERR_UNSUPPORTED, "group_mlp_max: tile does not fit shared memory");
    p.qt = tc.per_tile;
    dim3 grid(tc.tiles, d->nsets);
    p.nring = pick_ring(smem(tc.nb), p.total_chunks);
    const size_t bytes = smem(tc.nb) + (size_t)p.nring * CHUNK_BYTES;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err;
    if (tc.nb == 1) {
        if ((rc = set_smem(group_mlp_max_kernel<1>, bytes, "group_mlp_max smem"))) return rc;
        err = launch(group_mlp_max_kernel<1>, dim3(grid), dim3(LAUNCH_THREADS), bytes, st, p);
    } else {
        if ((rc = set_smem(group_mlp_max_kernel<2>, bytes, "group_mlp_max smem"))) return rc;
        err = launch(group_mlp_max_kernel<2>, dim3(grid), dim3(LAUNCH_THREADS), bytes, st, p);
    }
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "group_mlp_max launch");
}

extern "C" int elo_cost_volume_1(const elo_cost_volume_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume_1: null descriptor");
    int rc = check_window(&d->window_q, "cost_volume_1: bad window");
    if (rc) return rc;
    if (d->batch_size < 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || (d->C & 3) || !d->xyz1 || !d->xyz2 || !d->f1 ||
        !d->f2 || !d->weights_1 || !d->stage1_out || d->window_q.small_h != d->H || d->window_q.small_w != d->W ||
        d->window_q.stride_h != 1 || d->window_q.stride_w != 1)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume_1: bad arguments");
    if (d->window_q.K > 64) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_1: nsample_q > 64");
    if (d->batch_size == 0) return ELO_OK;
    Cv1Params p;
    static const int direct_gather = getenv("ELO_DIRECT_GATHER") ? atoi(getenv("ELO_DIRECT_GATHER")) : 1;
    p.direct_gather = direct_gather;
    p.sl_in_ring = 0;
    p.qs.H1 = d->H; p.qs.W1 = d->W; p.qs.oh = d->H; p.qs.ow = d->W; p.qs.qs_h = 1; p.qs.qs_w = 1;
    p.g = make_window(&d->window_q);
    p.total_q = (long long)d->batch_size * d->H * d->W;
    p.q_first = 0;
    if (d->query_begin != 0 || d->query_end != 0) {      // a sub-range of the pixels (row bands)
        if (d->query_begin < 0 || d->query_end > p.total_q || d->query_begin > d->query_end)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume: bad query range");
        p.q_first = d->query_begin;
        p.total_q = d->query_end;
        if (p.q_first == p.total_q) return ELO_OK;
    }
    p.C = d->C;
    const int xc = 10 + 2 * d->C;
    p.total_chunks = layer_chunks(xc, 128) + layer_chunks(128, 64) + layer_chunks(64, 64) + layer_chunks(10, 64) +
                     layer_chunks(128, 128) + layer_chunks(128, 64);
    p.xyz1 = d->xyz1; p.xyz2 = d->xyz2; p.f1 = d->f1; p.f2 = d->f2; p.random_hw = d->window_q.random_hw;
    p.weights = d->weights_1; p.out = d->stage1_out; p.dbg_nbr = d->dbg_nbr_q; p.nbr_in = d->nbr_q; p.tlog = g_tlog;
    const int kt = p.g.kt, xch = (xc + 3) & ~3;
    if (g_engine == 1) {
        if (!p.nbr_in) return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume_1: the tensor-core engine takes nbr_q from elo_multi_search");
        p.total_chunks = tc_layer_chunks(xc, 128) + tc_layer_chunks(128, 64) + tc_layer_chunks(64, 64) +
                         tc_layer_chunks(10, 64) + tc_layer_chunks(128, 128) + tc_layer_chunks(128, 64);
        if (p.g.K > 32) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_1: the tensor-core engine takes nsample_q <= 32");
        // two tiles per SM when the logits' pool staging can live in the weight ring (>= 3 slots hold 128 x 65 floats)
        size_t base = tc_base_smem(0, (size_t)(xch > POOL_LD64 ? xch : POOL_LD64) * TC_ROWS);
        p.nring = tc_pick_ring(base, p.total_chunks, /*two_per_sm=*/true);
        p.sl_in_ring = (p.nring >= 3 && base + (size_t)p.nring * TC_CHUNK_BYTES <= (size_t)SMEM_HALF) ? 1 : 0;
        if (!p.sl_in_ring) {
            base = tc_base_smem(0, (size_t)(xch > 2 * POOL_LD64 ? xch : 2 * POOL_LD64) * TC_ROWS);
            p.nring = tc_pick_ring(base, p.total_chunks);
        }
        if (p.nring < 2) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_1: tile does not fit shared memory");
        const TcChoice tc = choose_tc_tile(p.total_q - p.q_first, p.g.K, 1, p.sl_in_ring ? 2 : 1);
        p.qt = tc.per_tile;
        p.g.kt = 0;
        return launch_tc(cost_volume_1_tc_kernel, p, dim3(tc.tiles), base + (size_t)p.nring * TC_CHUNK_BYTES,
                         (cudaStream_t)stream, "cost_volume_1 (tensor core) launch");
    }
    auto smem = [&](int nb) {
        // the select-K scratch (8 warps x kt x 8 B) must fit in the A|Bf|Cb region it aliases
        if ((size_t)kt * 64 > (size_t)320 * 64 * nb * 4) return (size_t)1 << 30;
        return common_smem(kt, 64 * nb) + (size_t)(xch + 320) * 64 * nb * 4;
    };
    const TileChoice tc = choose_tile(p.total_q - p.q_first, p.g.K, 1, smem);
    if (tc.nb == 0) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_1: tile does not fit shared memory");
    p.qt = tc.per_tile;
    p.nring = pick_ring(smem(tc.nb), p.total_chunks);
    const size_t bytes = smem(tc.nb) + (size_t)p.nring * CHUNK_BYTES;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err;
    if (tc.nb == 1) {
        if ((rc = set_smem(cost_volume_1_kernel<1>, bytes, "cost_volume_1 smem"))) return rc;
        err = launch(cost_volume_1_kernel<1>, dim3(tc.tiles), dim3(LAUNCH_THREADS), bytes, st, p);
    } else {
        if ((rc = set_smem(cost_volume_1_kernel<2>, bytes, "cost_volume_1 smem"))) return rc;
        err = launch(cost_volume_1_kernel<2>, dim3(tc.tiles), dim3(LAUNCH_THREADS), bytes, st, p);
    }
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "cost_volume_1 launch");
}

extern "C" int elo_cost_volume_2(const elo_cost_volume_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume_2: null descriptor");
    int rc = check_window(&d->window_p, "cost_volume_2: bad window");
    if (rc) return rc;
    if (d->batch_size < 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || (d->C & 3) || !d->xyz1 || !d->f1 ||
        !d->weights_2 || !d->stage1_out || !d->out || d->window_p.small_h != d->H || d->window_p.small_w != d->W ||
        d->window_p.stride_h != 1 || d->window_p.stride_w != 1)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume_2: bad arguments");
    if (d->window_p.K > 64) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_2: nsample > 64");
    if (d->batch_size == 0) return ELO_OK;
    Cv2Params p;
    p.sl_in_ring = 0;
    p.qs.H1 = d->H; p.qs.W1 = d->W; p.qs.oh = d->H; p.qs.ow = d->W; p.qs.qs_h = 1; p.qs.qs_w = 1;
    p.g = make_window(&d->window_p);
    p.total_q = (long long)d->batch_size * d->H * d->W;
    p.q_first = 0;
    if (d->query_begin != 0 || d->query_end != 0) {      // a sub-range of the pixels (row bands)
        if (d->query_begin < 0 || d->query_end > p.total_q || d->query_begin > d->query_end)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume: bad query range");
        p.q_first = d->query_begin;
        p.total_q = d->query_end;
        if (p.q_first == p.total_q) return ELO_OK;
    }
    p.C = d->C;
    p.total_chunks = layer_chunks(10, 64) + layer_chunks(128 + d->C, 128) + layer_chunks(128, 64);
    p.xyz1 = d->xyz1; p.f1 = d->f1; p.cv1 = d->stage1_out; p.random_hw = d->window_p.random_hw;
    p.weights = d->weights_2; p.out = d->out; p.dbg_nbr = d->dbg_nbr_p; p.nbr_in = d->nbr_p; p.tlog = g_tlog;
    const int kt = p.g.kt;
    if (g_engine == 1) {
        if (!p.nbr_in) return set_error(ELO_ERR_INVALID_ARGUMENT, "cost_volume_2: the tensor-core engine takes nbr_p from elo_multi_search");
        p.total_chunks = tc_layer_chunks(10, 64) + tc_layer_chunks(128 + d->C, 128) + tc_layer_chunks(128, 64);
        if (p.g.K > 32) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_2: the tensor-core engine takes nsample <= 32");
        size_t base = tc_base_smem(0, (size_t)(80 + d->C) * TC_ROWS);
        p.nring = tc_pick_ring(base, p.total_chunks, /*two_per_sm=*/true);
        p.sl_in_ring = (p.nring >= 3 && base + (size_t)p.nring * TC_CHUNK_BYTES <= (size_t)SMEM_HALF) ? 1 : 0;
        if (!p.sl_in_ring) {
            base = tc_base_smem(0, (size_t)(80 + d->C + 2 * POOL_LD64) * TC_ROWS);
            p.nring = tc_pick_ring(base, p.total_chunks);
        }
        if (p.nring < 2) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_2: tile does not fit shared memory");
        const TcChoice tc = choose_tc_tile(p.total_q - p.q_first, p.g.K, 1, p.sl_in_ring ? 2 : 1);
        p.qt = tc.per_tile;
        p.g.kt = 0;
        return launch_tc(cost_volume_2_tc_kernel, p, dim3(tc.tiles), base + (size_t)p.nring * TC_CHUNK_BYTES,
                         (cudaStream_t)stream, "cost_volume_2 (tensor core) launch");
    }
    auto smem = [&](int nb) { return common_smem(kt, 64 * nb) + (size_t)(12 + 128 + d->C + 192) * 64 * nb * 4; };
    const TileChoice tc = choose_tile(p.total_q - p.q_first, p.g.K, 1, smem);
    if (tc.nb == 0) return set_error(ELO_ERR_UNSUPPORTED, "cost_volume_2: tile does not fit shared memory");
    p.qt = tc.per_tile;
    p.nring = pick_ring(smem(tc.nb), p.total_chunks);
    const size_t bytes = smem(tc.nb) + (size_t)p.nring * CHUNK_BYTES;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err;
    if (tc.nb == 1) {
        if ((rc = set_smem(cost_volume_2_kernel<1>, bytes, "cost_volume_2 smem"))) return rc;
        err = launch(cost_volume_2_kernel<1>, dim3(tc.tiles), dim3(LAUNCH_THREADS), bytes, st, p);
    } else {
        if ((rc = set_smem(cost_volume_2_kernel<2>, bytes, "cost_volume_2 smem"))) return rc;
        err = launch(cost_volume_2_kernel<2>, dim3(tc.tiles), dim3(LAUNCH_THREADS), bytes, st, p);
    }
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "cost_volume_2 launch");
}

extern "C" int elo_row_mlp(const elo_row_mlp_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: null descriptor");
    if (d->rows < 0 || d->nsets < 1 || d->nsets > 2 || d->num_phases < 1 || d->num_phases > 2)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: bad arguments");
    if (d->rows == 0) return ELO_OK;
    RowMlpParams p;
    p.tlog = g_tlog;
    p.rows = d->rows;
    p.nphase = d->num_phases;
    int chunks = 0, xmax = 0, prev_c = 0;
    for (int ph = 0; ph < 2; ++ph) {
        p.nsrc[ph] = 0; p.nl[ph] = 0;
        for (int i = 0; i < 3; ++i) { p.src_c[ph][i] = 0; p.src_prev[ph][i] = 0; p.cout[ph][i] = 0; }
        if (ph >= d->num_phases) continue;
        const elo_row_mlp_phase* f = &d->phase[ph];
        if (f->num_sources < 1 || f->num_sources > 3 || f->num_layers < 1 || f->num_layers > 3)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: bad phase");
        p.nsrc[ph] = f->num_sources; p.nl[ph] = f->num_layers;
        int cin = 0;
        for (int i = 0; i < f->num_sources; ++i) {
            p.src_c[ph][i] = f->channels[i];
            p.src_prev[ph][i] = f->from_previous[i] ? 1 : 0;
            if (f->channels[i] <= 0 || (f->channels[i] & 3)) return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: channels must be a positive multiple of 4");
            if (f->from_previous[i] && (ph == 0 || f->channels[i] != prev_c))
                return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: from_previous mismatch");
            cin += f->channels[i];
        }
        xmax = cin > xmax ? cin : xmax;
        for (int l = 0; l < f->num_layers; ++l) {
            if (!width_ok(f->cout[l])) return set_error(ELO_ERR_UNSUPPORTED, "row_mlp: layer widths must be 64 or 128");
            p.cout[ph][l] = f->cout[l];
            chunks += layer_chunks(cin, f->cout[l]);
            cin = f->cout[l];
        }
        prev_c = cin;
    }
    p.total_chunks = chunks;
    for (int s = 0; s < 2; ++s) {
        const int u = s < d->nsets ? s : 0;
        if (!d->weights[u] || !d->out[u]) return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: null pointer");
        p.weights[s] = d->weights[u]; p.out[s] = d->out[u]; p.out_phase0[s] = d->out_phase0[u];
        for (int ph = 0; ph < 2; ++ph)
            for (int i = 0; i < 3; ++i) {
                p.src[s][ph][i] = (ph < d->num_phases && i < d->phase[ph].num_sources) ? d->phase[ph].src[u][i] : nullptr;
                if (ph < d->num_phases && i < d->phase[ph].num_sources && !d->phase[ph].from_previous[i] && !p.src[s][ph][i])
                    return set_error(ELO_ERR_INVALID_ARGUMENT, "row_mlp: null source");
            }
    }
    const int xch = (xmax + 3) & ~3;
    if (g_engine == 1) {
        // staging slots: every distinct global source once (same pattern required for both sets)
        int stage_ch = 0, chunks_t = 0;
        for (int ph = 0; ph < p.nphase; ++ph) {
            int cin_t = 0;
            for (int i = 0; i < p.nsrc[ph]; ++i) {
                p.slot[ph][i] = -1; p.stage_here[ph][i] = 0;
                cin_t += p.src_c[ph][i];
                if (p.src_c[ph][i] & 15) return set_error(ELO_ERR_UNSUPPORTED, "row_mlp: tensor-core engine needs channels % 16 == 0");
                if (p.src_prev[ph][i]) continue;
                for (int ph2 = 0; ph2 <= ph && p.slot[ph][i] < 0; ++ph2)
                    for (int j = 0; j < (ph2 == ph ? i : p.nsrc[ph2]); ++j)
                        if (!p.src_prev[ph2][j] && p.src[0][ph2][j] == p.src[0][ph][i] && p.src[1][ph2][j] == p.src[1][ph][i] &&
                            p.src_c[ph2][j] == p.src_c[ph][i]) { p.slot[ph][i] = p.slot[ph2][j]; break; }
                if (p.slot[ph][i] < 0) { p.slot[ph][i] = stage_ch; p.stage_here[ph][i] = 1; stage_ch += p.src_c[ph][i]; }
            }
            if (cin_t > 192) return set_error(ELO_ERR_UNSUPPORTED, "row_mlp: more than 192 input channels");
            for (int l = 0; l < p.nl[ph]; ++l) { chunks_t += tc_layer_chunks(cin_t, p.cout[ph][l]); cin_t = p.cout[ph][l]; }
        }
        p.total_chunks = chunks_t;
        p.stage_ch = stage_ch;
        const size_t base = tc_base_smem(0, (size_t)stage_ch * TC_ROWS);
        // two tiles per SM where the staged sources leave room for a two-slot weight ring in half an SM's shared
        // memory (level 0: 144 staged channels): one tile's gather and stores then overlap the other's MMA chain
        static const bool two_ok = !(getenv("ELO_ROWMLP_TWO") && atoi(getenv("ELO_ROWMLP_TWO")) == 0);
        const bool two = two_ok && base + 2 * (size_t)TC_CHUNK_BYTES <= (size_t)SMEM_HALF;
        p.nring = tc_pick_ring(base, chunks_t, two);
        if (p.nring < 2) return set_error(ELO_ERR_UNSUPPORTED, "row_mlp: tile does not fit shared memory");
        const TcChoice tc = choose_tc_tile(p.rows, 1, d->nsets, two ? 2 : 1);
        p.rt = tc.per_tile;
        return launch_tc(row_mlp_tc_kernel, p, dim3(tc.tiles, d->nsets), base + (size_t)p.nring * TC_CHUNK_BYTES,
                         (cudaStream_t)stream, "row_mlp (tensor core) launch");
    }
    auto smem = [&](int nb) { return align16(2 * MAX_RING * 8) + (size_t)(xch + 256) * 64 * nb * 4; };
    const TileChoice tc = choose_tile(p.rows, 1, d->nsets, smem);
    if (tc.nb == 0) return set_error(ELO_ERR_UNSUPPORTED, "row_mlp: tile does not fit shared memory");
    p.rt = tc.per_tile;
    dim3 grid(tc.tiles, d->nsets);
    p.nring = pick_ring(smem(tc.nb), p.total_chunks);
    const size_t bytes = smem(tc.nb) + (size_t)p.nring * CHUNK_BYTES;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    cudaError_t err;
    if (tc.nb == 1) {
        if ((rc = set_smem(row_mlp_kernel<1>, bytes, "row_mlp smem"))) return rc;
        err = launch(row_mlp_kernel<1>, dim3(grid), dim3(LAUNCH_THREADS), bytes, st, p);
    } else {
        if ((rc = set_smem(row_mlp_kernel<2>, bytes, "row_mlp smem"))) return rc;
        err = launch(row_mlp_kernel<2>, dim3(grid), dim3(LAUNCH_THREADS), bytes, st, p);
    }
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "row_mlp launch");
}
