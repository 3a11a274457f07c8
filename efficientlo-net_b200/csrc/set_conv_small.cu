// set_conv_small.cu -- set-conv (utils/pointnet_util.py:179-250) for the narrow layers of the feature
// pyramid (pwclo_model.py:126-135: 6->8->8->16, 19->16->16->32, 35->32->32->64), sm_100a.
//
// These layers have hundreds of thousands of (query, neighbour) rows but only 240..4192 MACs per row,
// so a shared-memory GEMM tile would be all overhead.  Instead one warp owns 32/K queries: it runs the
// projection-aware random-K search for them (elo_search.cuh), then every lane takes ONE neighbour row,
// gathers its xyz/features straight from L2, pushes it through the three layers entirely in registers
// (weights are warp-uniform float4 broadcasts from shared memory) and the max over the K neighbours is
// a shuffle butterfly.  Nothing but the (B, n, C_out) result is written.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/elo_b200.h"
#include "elo_common.cuh"
#include "elo_tile.cuh"

namespace elo {

struct SmallParams {
    QuerySet qs;
    Window g;
    long long per_set;     // queries per parameter set (blockIdx.y); set s covers [q_base[s], q_base[s] + per_set)
    long long q_base[2];
    const float* xyz;      // (B, H, W, 3) query image == searched grid
    const float* feat;     // (B, H, W, CF) or null (all-zero features, pwclo_model.py:69-70)
    const int* random_hw[2];
    const float* weights;  // W1[(3+CF)][C1], b1[C1], W2[C1][C2], b2[C2], W3[C2][C3], b3[C3]  (BN folded)
    float* out;            // (B, n, C3)
    int* dbg_nbr;          // optional (B, n, K)
    const int* nbr_in;     // optional (B, n, K) from elo_multi_search: skip the search
};

template <int CIN, int COUT>
__device__ __forceinline__ void reg_layer(const float (&x)[CIN], float (&y)[COUT], const float* __restrict__ W,
                                          const float* __restrict__ b)
{
    static_assert(COUT % 4 == 0, "COUT must be a multiple of 4");
#pragma unroll
    for (int n = 0; n < COUT; n += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(b + n);
        y[n] = bv.x; y[n + 1] = bv.y; y[n + 2] = bv.z; y[n + 3] = bv.w;
    }
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
#pragma unroll
        for (int n = 0; n < COUT; n += 4) {
            const float4 w = *reinterpret_cast<const float4*>(W + k * COUT + n);
            y[n] = fmaf(x[k], w.x, y[n]);
            y[n + 1] = fmaf(x[k], w.y, y[n + 1]);
            y[n + 2] = fmaf(x[k], w.z, y[n + 2]);
            y[n + 3] = fmaf(x[k], w.w, y[n + 3]);
        }
    }
#pragma unroll
    for (int n = 0; n < COUT; ++n) y[n] = fmaxf(y[n], 0.f);
}

template <int CF, int C1, int C2, int C3, int K>
__global__ void __launch_bounds__(256) set_conv_small_kernel(const SmallParams p)
{
    constexpr int CIN = 3 + CF;
    constexpr int QPW = 32 / K;                      // queries per warp pass
    constexpr int NW = CIN * C1 + C1 + C1 * C2 + C2 + C2 * C3 + C3;
    static_assert(32 % K == 0, "K must divide 32");
    constexpr int NRES = (C3 + K - 1) / K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemCarver sc(smem_raw);
    float* wts = sc.take<float>(NW);
    int2* off = sc.take<int2>(p.g.kt);
    int* nbr_all = sc.take<int>(256);

    const Window g = p.g;
    pdl_trigger();
    for (int i = threadIdx.x; i < NW; i += blockDim.x) wts[i] = __ldg(p.weights + i);     // constants: before the wait
    pdl_wait();
    build_offsets(off, p.random_hw[blockIdx.y], g.kt, g.kH, g.kW, blockDim.x);
    __syncthreads();
    const float* W1 = wts;
    const float* b1 = W1 + CIN * C1;
    const float* W2 = b1 + C1;
    const float* b2 = W2 + C1 * C2;
    const float* W3 = b2 + C2;
    const float* b3 = W3 + C2 * C3;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* nbr = nbr_all + warp * 32;
    const int nq = p.qs.oh * p.qs.ow, cells = g.h2 * g.w2;
    const long long groups = (p.per_set + QPW - 1) / QPW;
    const long long q_base = p.q_base[blockIdx.y], q_end = q_base + p.per_set;
    const int sub = lane / K, kk = lane % K;

    const int nwarps = blockDim.x >> 5;
    for (long long grp = (long long)blockIdx.x * nwarps + warp; grp < groups; grp += (long long)gridDim.x * nwarps) {
        nbr[lane] = -1;
        float cx[QPW], cy[QPW], cz[QPW];
        int cb[QPW];
        __syncwarp();
#pragma unroll
        for (int qi = 0; qi < QPW; ++qi) {
            const long long gq = q_base + grp * QPW + qi;
            cx[qi] = cy[qi] = cz[qi] = 0.f;
            cb[qi] = -1;
            if (gq >= q_end) continue;
            const int b = (int)(gq / nq);
            int h, w;
            query_cell(p.qs, (int)(gq % nq), h, w);
            const float* c = p.xyz + ((size_t)b * p.qs.H1 * p.qs.W1 + (size_t)h * p.qs.W1 + w) * 3;
            cx[qi] = __ldg(c); cy[qi] = __ldg(c + 1); cz[qi] = __ldg(c + 2);
            cb[qi] = b;
            if (p.nbr_in != nullptr) {
                if (lane < K) nbr[qi * K + lane] = __ldg(p.nbr_in + gq * K + lane);
                continue;
            }
            if (fmaxf(sq3(cx[qi], cy[qi], cz[qi]), 1e-10f) <= 1e-10f) continue;     // empty centre: all masked
            int* row = nbr + qi * K;
            auto emit = [&](int slot, int hh, int ww) { row[slot] = hh * g.w2 + ww; };
            search_random_k(p.xyz + (size_t)b * cells * 3, off, g, h / g.stride_h, w / g.stride_w, cx[qi], cy[qi],
                            cz[qi], emit);
        }
        __syncwarp();

        // this lane's row: neighbour kk of query `sub`
        float px = 0.f, py = 0.f, pz = 0.f;
        int b = -1;
#pragma unroll
        for (int qi = 0; qi < QPW; ++qi)
            if (qi == sub) { px = cx[qi]; py = cy[qi]; pz = cz[qi]; b = cb[qi]; }
        const int cell = nbr[lane];
        const bool valid = cell >= 0;
        float x[CIN];
        {
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (valid) {
                const float* s = p.xyz + ((size_t)b * cells + cell) * 3;
                qx = __ldg(s); qy = __ldg(s + 1); qz = __ldg(s + 2);
            }
            x[0] = qx - px; x[1] = qy - py; x[2] = qz - pz;
#pragma unroll
            for (int c = 0; c < CF; ++c) x[3 + c] = 0.f;
            if (valid && p.feat != nullptr) {
                const float* f = p.feat + ((size_t)b * cells + cell) * CF;
                if (CF % 4 == 0) {
#pragma unroll
                    for (int c = 0; c < CF / 4; ++c) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(f) + c);
                        x[3 + 4 * c] = v.x; x[4 + 4 * c] = v.y; x[5 + 4 * c] = v.z; x[6 + 4 * c] = v.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < CF; ++c) x[3 + c] = __ldg(f + c);
                }
            }
        }
        float h1[C1], h2[C2], y[C3];
        reg_layer<CIN, C1>(x, h1, W1, b1);
        reg_layer<C1, C2>(h1, h2, W2, b2);
        reg_layer<C2, C3>(h2, y, W3, b3);

        // (y * mask), then max over the K lanes of the query
        float res[NRES];
#pragma unroll
        for (int j = 0; j < NRES; ++j) res[j] = 0.f;
#pragma unroll
        for (int c = 0; c < C3; ++c) {
            float v = valid ? y[c] : 0.f;
#pragma unroll
            for (int o = K / 2; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
            if (kk == (c % K)) res[c / K] = v;
        }
        const long long gq = q_base + grp * QPW + sub;
        if (gq < q_end) {
#pragma unroll
            for (int j = 0; j < NRES; ++j)
                if (j * K + kk < C3) p.out[gq * C3 + j * K + kk] = res[j];
            if (p.dbg_nbr != nullptr) p.dbg_nbr[gq * K + kk] = cell;
        }
        __syncwarp();
    }
}

template <int CF, int C1, int C2, int C3, int K>
static int launch_small(const SmallParams& p, int nsets, cudaStream_t st)
{
    constexpr int NW = (3 + CF) * C1 + C1 + C1 * C2 + C2 + C2 * C3 + C3;
    const size_t smem = ((NW * 4 + 15) & ~15) + (((size_t)p.g.kt * 8 + 15) & ~15) + 1024;
    constexpr int QPW = 32 / K;
    const long long groups = (p.per_set + QPW - 1) / QPW;
    // few queries: fewer warps per CTA so that the work spreads over all SMs (each CTA re-reads the
    // <= 17 KB of weights from L2, which is cheap); many queries: 8 warps, grid-stride
    const long long sms = device_info().sm_count;
    long long wpc = (groups * nsets + 2 * sms - 1) / (2 * sms);
    wpc = wpc < 4 ? 4 : (wpc > 8 ? 8 : wpc);      // at least 4 warps share a CTA's weight load
    long long ctas = (groups + wpc - 1) / wpc;
    static const int cap_per_sm = getenv("ELO_SETCONV_CAP") ? atoi(getenv("ELO_SETCONV_CAP")) : 4;
    const long long cap = sms * cap_per_sm / nsets;
    if (ctas > cap) ctas = cap;
    auto kern = set_conv_small_kernel<CF, C1, C2, C3, K>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, "set_conv_small smem");
    }
    cudaError_t err = launch(kern, dim3((unsigned)ctas, nsets), dim3((unsigned)wpc * 32), smem, st, p);
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "set_conv_small launch");
}

}  // namespace elo

using namespace elo;

extern "C" int elo_set_conv_small(const elo_group_mlp_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "set_conv_small: null descriptor");
    const elo_window* w = &d->window[0];
    if (w->kernel_size_H <= 0 || w->kernel_size_W <= 0 || w->K <= 0 || !(w->distance > 0) || w->stride_h <= 0 ||
        w->stride_w <= 0 || !w->random_hw || d->batch_size < 0 || d->nsets < 1 || d->nsets > 2 || d->num_layers != 3 || !d->xyz1 ||
        !d->weights[0] || !d->out[0] || d->queries.H <= 0 || d->queries.W <= 0 || d->queries.out_h <= 0 ||
        d->queries.out_w <= 0 || d->queries.q_stride_h <= 0 || d->queries.q_stride_w <= 0 ||
        (d->queries.out_h - 1) * d->queries.q_stride_h >= d->queries.H ||
        (d->queries.out_w - 1) * d->queries.q_stride_w >= d->queries.W)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "set_conv_small: bad arguments");
    if (d->xyz2 != d->xyz1 || w->small_h != d->queries.H || w->small_w != d->queries.W)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "set_conv_small: the searched grid must be the query image");
    if ((long long)w->kernel_size_H * w->kernel_size_W > 5000 || d->queries.H >= 32768 || d->queries.W >= 32768)
        return set_error(ELO_ERR_UNSUPPORTED, "set_conv_small: window or grid too large");
    if (d->batch_size == 0) return ELO_OK;
    SmallParams p;
    p.qs.H1 = d->queries.H; p.qs.W1 = d->queries.W; p.qs.oh = d->queries.out_h; p.qs.ow = d->queries.out_w;
    p.qs.qs_h = d->queries.q_stride_h; p.qs.qs_w = d->queries.q_stride_w;
    p.g.h2 = w->small_h; p.g.w2 = w->small_w; p.g.kH = w->kernel_size_H; p.g.kW = w->kernel_size_W;
    p.g.kt = w->kernel_size_H * w->kernel_size_W; p.g.stride_h = w->stride_h; p.g.stride_w = w->stride_w;
    p.g.K = w->K; p.g.flag_copy = 0; p.g.d2max = w->distance * w->distance;
    p.per_set = (long long)d->batch_size * p.qs.oh * p.qs.ow;
    long long qb = 0;                        // optional sub-range of each set's queries (row bands)
    if (d->query_begin != 0 || d->query_end != 0) {
        if (d->query_begin < 0 || d->query_end > p.per_set || d->query_begin > d->query_end)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "set_conv_small: bad query range");
        qb = d->query_begin;
        p.per_set = d->query_end - d->query_begin;
        if (p.per_set == 0) return ELO_OK;
    }
    for (int s = 0; s < 2; ++s) {
        const int u = s < d->nsets ? s : 0;
        if (!d->window[u].random_hw || d->set_batch_offset[u] < 0)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "set_conv_small: bad parameter set");
        p.random_hw[s] = d->window[u].random_hw;
        p.q_base[s] = (long long)d->set_batch_offset[u] * p.qs.oh * p.qs.ow + qb;
    }
    p.xyz = d->xyz1; p.feat = d->feat2[0]; p.weights = d->weights[0];
    p.out = d->out[0]; p.dbg_nbr = d->dbg_nbr[0]; p.nbr_in = d->nbr[0];
    cudaStream_t st = (cudaStream_t)stream;
    const int cf = d->feat_channels, c1 = d->cout[0], c2 = d->cout[1], c3 = d->cout[2], K = w->K;
    if (cf == 3 && c1 == 8 && c2 == 8 && c3 == 16 && K == 32) return launch_small<3, 8, 8, 16, 32>(p, d->nsets, st);
    if (cf == 16 && c1 == 16 && c2 == 16 && c3 == 32 && K == 32) return launch_small<16, 16, 16, 32, 32>(p, d->nsets, st);
    if (cf == 32 && c1 == 32 && c2 == 32 && c3 == 64 && K == 16) return launch_small<32, 32, 32, 64, 16>(p, d->nsets, st);
    return set_error(ELO_ERR_UNSUPPORTED,
                     "set_conv_small: only the pyramid's (3;8,8,16;K32), (16;16,16,32;K32), (32;32,32,64;K16) layers");
}
