// fused_conv_index.cu -- stand-alone projection-aware neighbour-search ops for sm_100a.
//
// Drop-in for the reference's native seam (file:line relative to /root/reference):
//   FusedConvSelectKLauncher  tf_ops/2d_conv_select_k/fused_conv_g.cu:215  (kernel :11-209)
//   FusedConvRandomKLauncher  tf_ops/2d_conv_random_k/fused_conv_g.cu:162  (kernel :13-156)
// and for the zero-fill the op wrapper does before launching (fused_conv.cpp:154-166).
//
// Design (vs the reference's <<<B,256>>>, one thread per query, 60 KB of local memory per thread):
//   * one warp per query, grid over all B*npoints queries (persistent, grid-stride), so the whole
//     chip is busy at B = 1;
//   * the window's scan-order offset table is built once per CTA in shared memory;
//   * lanes test 32 window cells per step; slots and run-length counts come from ballots;
//   * every output element is written exactly once with lane-contiguous (coalesced) stores --
//     the kt-wide valid_* rows dominate the op's bytes, so the op is HBM-write bound
//     (kernel K1) -- and no memset pass is needed.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/elo_b200.h"
#include "elo_common.cuh"
#include "elo_search.cuh"
#include "elo_tile.cuh"

namespace elo {

struct IndexParams {
    int B, H, W, N;
    Window g;
    const float* xyz1;
    const float* xyz2;
    const int* idx_n2;
    const int* random_hw;
    int* out_idx;      // (B, N, K, 3)
    float* out_valid;  // (B, N, kt) or null
    float* out_vdis;   // (B, N, kt) or null
    float* out_mask;   // (B, N, K)
    long long total;   // B * N
    bool fast_ok;      // register fast path allowed (K <= 32, distance^2 < 1e10)
};

__device__ __forceinline__ void fill_counts(float* row, int kt, int ones, int lane)
{
    if (row == nullptr) return;
    for (int i = lane; i < kt; i += 32) row[i] = i < ones ? 1.0f : 0.0f;
}

// slots [from, K): either all zero, or (b, hh, ww) with mask 1 (flag_copy)
__device__ __forceinline__ void fill_slots(int* o_idx, float* o_mask, int from, int K, bool copy,
                                           int b, int packed, int lane)
{
    const int hh = packed >> 16, ww = packed & 0xffff;
    for (int k = from + lane; k < K; k += 32) {
        o_idx[k * 3 + 0] = copy ? b : 0;
        o_idx[k * 3 + 1] = copy ? hh : 0;
        o_idx[k * 3 + 2] = copy ? ww : 0;
        o_mask[k] = copy ? 1.0f : 0.0f;
    }
}

template <bool SELECT>
__global__ void __launch_bounds__(256) fused_conv_index_kernel(const IndexParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();
    const Window g = p.g;
    int2* off = reinterpret_cast<int2*>(smem_raw);
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* dist = nullptr;
    int* hw = nullptr;
    if (SELECT) {
        float* dist_all = reinterpret_cast<float*>(off + g.kt);
        int* hw_all = reinterpret_cast<int*>(dist_all + (size_t)nwarps * g.kt);
        dist = dist_all + (size_t)warp * g.kt;
        hw = hw_all + (size_t)warp * g.kt;
    }
    build_offsets(off, p.random_hw, g.kt, g.kH, g.kW, blockDim.x);
    __syncthreads();

    for (long long q = (long long)blockIdx.x * nwarps + warp; q < p.total;
         q += (long long)gridDim.x * nwarps) {
        const int b = (int)(q / p.N);
        const int h = __ldg(p.idx_n2 + q * 2), w = __ldg(p.idx_n2 + q * 2 + 1);
        int* o_idx = p.out_idx + q * g.K * 3;
        float* o_mask = p.out_mask + q * g.K;
        float* o_valid = p.out_valid ? p.out_valid + q * g.kt : nullptr;
        float* o_vdis = p.out_vdis ? p.out_vdis + q * g.kt : nullptr;

        float xc = 0.f, yc = 0.f, zc = 0.f;
        const bool inside = h >= 0 && h < p.H && w >= 0 && w < p.W;  // reference: out-of-range = UB
        if (inside) {
            const float* c = p.xyz1 + ((size_t)b * p.H * p.W + (size_t)h * p.W + w) * 3;
            xc = __ldg(c); yc = __ldg(c + 1); zc = __ldg(c + 2);
        }
        // invalid centre (empty pixel): the whole row stays zero (reference :61-69)
        if (!inside || fmaxf(sq3(xc, yc, zc), 1e-10f) <= 1e-10f) {
            fill_slots(o_idx, o_mask, 0, g.K, false, 0, 0, lane);
            fill_counts(o_valid, g.kt, 0, lane);
            fill_counts(o_vdis, g.kt, 0, lane);
            continue;
        }
        const float* g2 = p.xyz2 + (size_t)b * g.h2 * g.w2 * 3;
        const int ch = h / g.stride_h, cw = w / g.stride_w;
        auto emit = [&](int slot, int hh, int ww) {
            o_idx[slot * 3 + 0] = b;
            o_idx[slot * 3 + 1] = hh;
            o_idx[slot * 3 + 2] = ww;
            o_mask[slot] = 1.0f;
        };
        int filled;
        const SearchCounts c = search_query<SELECT>(g2, off, g, ch, cw, xc, yc, zc, dist, hw, &filled, p.fast_ok, emit);
        // select-K duplicates entry 0 even when nothing was in range (mask 1, index (b,0,0));
        // random-K only once a first neighbour was accepted (reference select :180-192, random :126-138)
        const bool copy = g.flag_copy == 1 && (SELECT || c.nsel > 0);
        fill_slots(o_idx, o_mask, filled, g.K, copy, b, c.first, lane);
        fill_counts(o_valid, g.kt, c.nvalid, lane);
        fill_counts(o_vdis, g.kt, c.nsel, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Batched search: up to ELO_MAX_SEARCH independent neighbour searches in ONE launch, each writing the
// compact table the fused blocks consume: nbr[b, n, k] = linear cell (hh * small_w + ww) of the k-th
// selected neighbour in the searched grid, -1 where the reference's mask is 0.  Every search of the
// network that depends on the same xyz grids runs together (all 11 of the un-warped pyramid; the 4 of
// a refinement level), thousands of warps in flight instead of a few per GEMM tile.
struct SearchSpec {
    QuerySet qs;
    Window g;
    int select, fast_ok;
    long long q_first;        // queries [q_first, total_q) of the batch * oh * ow are searched
    long long total_q;
    int cta_begin;            // first CTA of this spec
    const float* xyz1;
    const float* xyz2;
    const int* random_hw;
    int* out_nbr;
};

struct MultiSearchParams {
    int nspec;
    SearchSpec spec[ELO_MAX_SEARCH];
};

__global__ void __launch_bounds__(256) multi_search_kernel(const __grid_constant__ MultiSearchParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();          // the grids (and, in eager calls, the scan orders) come from the kernels before
    int si = 0;
    while (si + 1 < p.nspec && (int)blockIdx.x >= p.spec[si + 1].cta_begin) ++si;
    const SearchSpec& sp = p.spec[si];
    const Window g = sp.g;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int2* off = reinterpret_cast<int2*>(smem_raw);
    float* dist = reinterpret_cast<float*>(off + g.kt) + (size_t)warp * g.kt;
    int* hw = reinterpret_cast<int*>(reinterpret_cast<float*>(off + g.kt) + (size_t)nwarps * g.kt) + (size_t)warp * g.kt;
    build_offsets(off, sp.random_hw, g.kt, g.kH, g.kW, blockDim.x);
    __syncthreads();
    const int nq = sp.qs.oh * sp.qs.ow;
    const int cta_end = (si + 1 < p.nspec) ? p.spec[si + 1].cta_begin : (int)gridDim.x;
    const int ctas = cta_end - sp.cta_begin;
    for (long long q = sp.q_first + (long long)((int)blockIdx.x - sp.cta_begin) * nwarps + warp; q < sp.total_q;
         q += (long long)ctas * nwarps) {
        int* row = sp.out_nbr + q * g.K;
        for (int k = lane; k < g.K; k += 32) row[k] = -1;
        const int b = (int)(q / nq);
        int h, w;
        query_cell(sp.qs, (int)(q % nq), h, w);
        const float* c = sp.xyz1 + ((size_t)b * sp.qs.H1 * sp.qs.W1 + (size_t)h * sp.qs.W1 + w) * 3;
        const float xc = __ldg(c), yc = __ldg(c + 1), zc = __ldg(c + 2);
        __syncwarp();
        if (fmaxf(sq3(xc, yc, zc), 1e-10f) <= 1e-10f) continue;
        const float* g2 = sp.xyz2 + (size_t)b * g.h2 * g.w2 * 3;
        auto emit = [&](int slot, int hh, int ww) { row[slot] = hh * g.w2 + ww; };
        int written;
        if (sp.select)
            search_query<true>(g2, off, g, h / g.stride_h, w / g.stride_w, xc, yc, zc, dist, hw, &written, sp.fast_ok != 0, emit);
        else
            search_query<false>(g2, off, g, h / g.stride_h, w / g.stride_w, xc, yc, zc, dist, hw, &written, sp.fast_ok != 0, emit);
        __syncwarp();
    }
}

// fused_conv_tiled.cu: thread-per-query kernel over a staged tile, for calls with many queries
int launch_index_tiled(bool select, int B, int H, int W, int N, const Window& g, const float* xyz1, const float* xyz2,
                       const int* idx_n2, const int* random_hw, int* out_idx, float* out_valid, float* out_vdis,
                       float* out_mask, cudaStream_t stream, int* rc);

int launch_search_tiled(bool select, int B, int H1, int W1, int oh, int ow, int qs_h, int qs_w, const Window& g,
                        const float* xyz1, const float* xyz2, const int* random_hw, int* out_nbr, cudaStream_t stream,
                        int* rc);

static int launch_index(bool select, int B, int H, int W, int N, int kH, int kW, int K, int flag_copy,
                        float distance, int stride_h, int stride_w, const float* xyz1,
                        const float* xyz2, const int* idx_n2, const int* random_hw, int* out_idx,
                        float* out_valid, float* out_vdis, float* out_mask, int h2, int w2,
                        cudaStream_t stream)
{
    // fused_conv.cpp:79-100
    if (N <= 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive npoints");
    if (kH <= 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive kernel_size_H");
    if (kW <= 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive kernel_size_W");
    if (K <= 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive K");
    if (flag_copy <= -1) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects 0 OR 1 flag_copy");
    if (!(distance > 0)) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive distance");
    if (stride_h <= 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive stride_h");
    if (stride_w <= 0) return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive stride_w");
    if (B < 0 || H <= 0 || W <= 0 || h2 <= 0 || w2 <= 0)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv expects positive grid extents");
    if (!xyz1 || !xyz2 || !idx_n2 || !random_hw || !out_idx || !out_mask)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "FusedConv: null pointer");
    const long long kt = (long long)kH * kW;
    if (kt > 5000 || K > 5000)
        return set_error(ELO_ERR_UNSUPPORTED, "FusedConv: kernel_size_H*kernel_size_W and K are limited to 5000");
    if (H >= 32768 || W >= 32768 || h2 >= 32768 || w2 >= 32768)
        return set_error(ELO_ERR_UNSUPPORTED, "FusedConv: grid extents are limited to 32767");
    if (B == 0) return ELO_OK;

    IndexParams p;
    p.B = B; p.H = H; p.W = W; p.N = N;
    p.g.h2 = h2; p.g.w2 = w2; p.g.kH = kH; p.g.kW = kW; p.g.kt = (int)kt;
    p.g.stride_h = stride_h; p.g.stride_w = stride_w; p.g.K = K; p.g.flag_copy = flag_copy;
    p.g.d2max = distance * distance;
    p.xyz1 = xyz1; p.xyz2 = xyz2; p.idx_n2 = idx_n2; p.random_hw = random_hw;
    p.out_idx = out_idx; p.out_valid = out_valid; p.out_vdis = out_vdis; p.out_mask = out_mask;
    p.total = (long long)B * N;
    p.fast_ok = K <= 32 && p.g.d2max < 1e10f;

    {
        int rc = ELO_OK;
        if (launch_index_tiled(select, B, H, W, N, p.g, xyz1, xyz2, idx_n2, random_hw, out_idx, out_valid, out_vdis,
                               out_mask, stream, &rc))
            return rc;
    }

    int warps = 8;
    size_t smem = (size_t)kt * sizeof(int2);
    if (select) {
        const size_t per_warp = (size_t)kt * 8;
        const size_t budget = 200 * 1024 - smem;
        while (warps > 1 && per_warp * warps > budget) warps >>= 1;
        smem += per_warp * warps;
    }
    const DeviceInfo& dev = device_info();
    long long ctas = (p.total + warps - 1) / warps;
    const long long resident = (long long)dev.sm_count * (select ? 4 : 8);
    if (ctas > resident) ctas = resident;

    cudaError_t err;
    if (select) {
        if (smem > 48 * 1024) {
            err = cudaFuncSetAttribute(fused_conv_index_kernel<true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return set_cuda_error(err, "cudaFuncSetAttribute(select_k)");
        }
        err = launch(fused_conv_index_kernel<true>, dim3((unsigned)ctas), dim3(warps * 32), smem, stream, p);
    } else {
        err = launch(fused_conv_index_kernel<false>, dim3((unsigned)ctas), dim3(warps * 32), smem, stream, p);
    }
    if (err != cudaSuccess) return set_cuda_error(err, select ? "fused_conv_select_k launch" : "fused_conv_random_k launch");
    return ELO_OK;
}

}  // namespace elo

extern "C" int elo_multi_search(const elo_search_desc* specs, int nspec, void* stream)
{
    using namespace elo;
    if (!specs || nspec < 1 || nspec > ELO_MAX_SEARCH)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "multi_search: 1..ELO_MAX_SEARCH specs");
    MultiSearchParams p;
    p.nspec = 0;
    const int sms = device_info().sm_count;
    const int warps = 8;
    size_t smem = 0;
    long long ctas_total = 0;
    for (int i = 0; i < nspec; ++i) {
        const elo_search_desc* d = &specs[i];
        const elo_window* w = &d->window;
        if (w->kernel_size_H <= 0 || w->kernel_size_W <= 0 || w->K <= 0 || !(w->distance > 0) || w->stride_h <= 0 ||
            w->stride_w <= 0 || w->small_h <= 0 || w->small_w <= 0 || !w->random_hw || !d->xyz1 || !d->xyz2 ||
            !d->out_nbr || d->batch_size < 0 || d->queries.H <= 0 || d->queries.W <= 0 || d->queries.out_h <= 0 ||
            d->queries.out_w <= 0 || d->queries.q_stride_h <= 0 || d->queries.q_stride_w <= 0 ||
            (d->queries.out_h - 1) * d->queries.q_stride_h >= d->queries.H ||
            (d->queries.out_w - 1) * d->queries.q_stride_w >= d->queries.W)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "multi_search: bad spec");
        const long long kt = (long long)w->kernel_size_H * w->kernel_size_W;
        if (kt > 5000 || w->K > 5000 || w->small_h >= 32768 || w->small_w >= 32768)
            return set_error(ELO_ERR_UNSUPPORTED, "multi_search: window / K limited to 5000, grids to 32767");
        if (d->batch_size == 0) continue;
        // Experiment (ELO_SEARCH_TILED=1, off by default): under the throughput tile policy, dense-query searches
        // with wide windows on the tile-staged thread-per-query kernel (its own launch).  Measured: less SM time
        // per search, but 10 more launches per forward and 25-40 us single-search latencies that stretch a forward
        // from 0.54 to 0.78 ms -- 5224 instead of 6409 pairs/s with 12 forwards in flight.
        static const bool search_tiled = getenv("ELO_SEARCH_TILED") != nullptr && getenv("ELO_SEARCH_TILED")[0] == '1';
        if (search_tiled && elo_get_tile_policy() == 1 && kt >= 64 && d->queries.q_stride_h == 1 && d->queries.q_stride_w == 1) {
            Window gw;
            gw.h2 = w->small_h; gw.w2 = w->small_w; gw.kH = w->kernel_size_H; gw.kW = w->kernel_size_W; gw.kt = (int)kt;
            gw.stride_h = w->stride_h; gw.stride_w = w->stride_w; gw.K = w->K; gw.flag_copy = 0;
            gw.d2max = w->distance * w->distance;
            int rc = ELO_OK;
            if (launch_search_tiled(d->select != 0, d->batch_size, d->queries.H, d->queries.W, d->queries.out_h,
                                    d->queries.out_w, 1, 1, gw, d->xyz1, d->xyz2, w->random_hw, d->out_nbr,
                                    (cudaStream_t)stream, &rc)) {
                if (rc != ELO_OK) return rc;
                continue;
            }
        }
        SearchSpec& sp = p.spec[p.nspec++];
        sp.qs.H1 = d->queries.H; sp.qs.W1 = d->queries.W; sp.qs.oh = d->queries.out_h; sp.qs.ow = d->queries.out_w;
        sp.qs.qs_h = d->queries.q_stride_h; sp.qs.qs_w = d->queries.q_stride_w;
        sp.g.h2 = w->small_h; sp.g.w2 = w->small_w; sp.g.kH = w->kernel_size_H; sp.g.kW = w->kernel_size_W;
        sp.g.kt = (int)kt; sp.g.stride_h = w->stride_h; sp.g.stride_w = w->stride_w; sp.g.K = w->K;
        sp.g.flag_copy = 0; sp.g.d2max = w->distance * w->distance;
        sp.select = d->select ? 1 : 0;
        sp.fast_ok = (w->K <= 32 && sp.g.d2max < 1e10f) ? 1 : 0;
        sp.total_q = (long long)d->batch_size * sp.qs.oh * sp.qs.ow;
        sp.q_first = 0;
        if (d->query_begin != 0 || d->query_end != 0) {
            if (d->query_begin < 0 || d->query_end > sp.total_q || d->query_begin > d->query_end)
                return set_error(ELO_ERR_INVALID_ARGUMENT, "multi_search: bad query range");
            sp.q_first = d->query_begin;
            sp.total_q = d->query_end;
            if (sp.q_first == sp.total_q) { --p.nspec; continue; }
        }
        sp.xyz1 = d->xyz1; sp.xyz2 = d->xyz2; sp.random_hw = w->random_hw; sp.out_nbr = d->out_nbr;
        long long ctas = (sp.total_q - sp.q_first + warps - 1) / warps;
        // A CTA pays a fixed cost (scan-order table, barrier, the first round trip) for its 8 warps.  Latency policy:
        // a warp per query, up to 8 CTAs per SM.  Throughput policy (several forwards in flight): up to four queries per
        // warp but at least one CTA per SM -- the fixed cost is amortised and the search leaves more of the SMs to the
        // other forwards (measured, 12 forwards in flight: 6 401 -> 6 616 pairs/s; B = 8: unchanged).
        static const int cap_per_sm = getenv("ELO_SEARCH_CAP") ? atoi(getenv("ELO_SEARCH_CAP")) : 0;
        long long cap = (long long)sms * (cap_per_sm > 0 ? cap_per_sm : 8);
        if (cap_per_sm <= 0 && elo_get_tile_policy() == 1) {
            const long long four = (sp.total_q - sp.q_first + warps * 4 - 1) / (warps * 4);
            cap = four > sms ? four : sms;
        }
        if (ctas > cap) ctas = cap;
        sp.cta_begin = (int)ctas_total;
        ctas_total += ctas;
        const size_t need = (size_t)kt * 8 + (d->select ? (size_t)kt * 8 * warps : 0);
        if (need > smem) smem = need;
    }
    if (p.nspec == 0) return ELO_OK;
    if (smem > 200 * 1024) return set_error(ELO_ERR_UNSUPPORTED, "multi_search: window too large for one CTA");
    cudaError_t err;
    if (smem > 48 * 1024) {
        err = cudaFuncSetAttribute(multi_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return set_cuda_error(err, "cudaFuncSetAttribute(multi_search)");
    }
    err = launch(multi_search_kernel, dim3((unsigned)ctas_total), dim3(warps * 32), smem, (cudaStream_t)stream, p);
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "multi_search launch");
}

extern "C" int elo_fused_conv_select_k(int batch_size, int H, int W, int npoints, int kernel_size_H,
                                       int kernel_size_W, int K, int flag_copy, float distance,
                                       int stride_h, int stride_w, const float* xyz1,
                                       const float* xyz2, const int* idx_n2, const int* random_hw,
                                       int* selected_bhw_idx, float* valid_idx,
                                       float* valid_in_dis_idx, float* selected_mask, int small_h,
                                       int small_w, void* stream)
{
    return elo::launch_index(true, batch_size, H, W, npoints, kernel_size_H, kernel_size_W, K,
                             flag_copy, distance, stride_h, stride_w, xyz1, xyz2, idx_n2, random_hw,
                             selected_bhw_idx, valid_idx, valid_in_dis_idx, selected_mask, small_h,
                             small_w, (cudaStream_t)stream);
}

extern "C" int elo_fused_conv_random_k(int batch_size, int H, int W, int npoints, int kernel_size_H,
                                       int kernel_size_W, int K, int flag_copy, float distance,
                                       int stride_h, int stride_w, const float* xyz1,
                                       const float* xyz2, const int* idx_n2, const int* random_hw,
                                       int* selected_bhw_idx, float* valid_idx,
                                       float* valid_in_dis_idx, float* selected_mask, int small_h,
                                       int small_w, void* stream)
{
    return elo::launch_index(false, batch_size, H, W, npoints, kernel_size_H, kernel_size_W, K,
                             flag_copy, distance, stride_h, stride_w, xyz1, xyz2, idx_n2, random_hw,
                             selected_bhw_idx, valid_idx, valid_in_dis_idx, selected_mask, small_h,
                             small_w, (cudaStream_t)stream);
}
