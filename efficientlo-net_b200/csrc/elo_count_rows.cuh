// elo_count_rows.cuh -- writer of the two count tensors of the index ops (valid_idx, valid_in_dis_idx).
//
// A row of either tensor is `count` ones followed by kt - count zeros (run-length counts, not per-position flags:
// tf_ops/2d_conv_select_k/fused_conv_g.cu:106-139).  They are 83 % of the op's bytes (2 kt floats per query), so the
// writer is built for the store path: a warp writes the 32 rows of its own queries with 16-byte streaming stores;
// four rows are kt float4s, 16-byte aligned; a lane keeps the same float4 column for every block of four rows, so which
// rows its four elements belong to (at most two when kt >= 4) and their positions are loop-invariant, and an
// element is saturate(count - position): one FADD.SAT on the FP32 pipe per element and row.
// Used by the query warps and by the store warp of the tile-staged index kernel (fused_conv_tiled.cu).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "elo_search.cuh"

namespace elo {

__device__ __forceinline__ unsigned udiv_magic(unsigned e, unsigned magic, unsigned d, unsigned& rem)
{
    unsigned q = __umulhi(e, magic);
    rem = e - q * d;
    if (rem >= d) { rem -= d; ++q; }      // magic = 2^32 - 1 stands in for d = 1
    return q;
}

// Rows [0, nrows) written by ONE warp: o_valid / o_vdis point at the first row (either may be NULL; the first row
// must start on a 16-byte boundary for the vector path, i.e. its index is a multiple of 4 or kt is), nv / ns at the
// first row's count in shared memory (floats).  vec_ok: both tensors are 16-byte aligned.  nrows is 32 for a query
// warp that writes its own rows and the CTA's whole row count for the store warp of fused_conv_tiled.cu.
__device__ __forceinline__ void write_count_rows(float* o_valid, float* o_vdis, int nrows, int kt, bool vec_ok,
                                                 unsigned magic_kt, const float* nv, const float* ns, int lane)
{
    if ((o_valid == nullptr && o_vdis == nullptr) || nrows <= 0) return;
    const int nblk = (vec_ok && kt >= 4) ? nrows / 4 : 0;
    const int nfi = nblk > 0 ? (kt + 31) / 32 : 0;         // column chunks of 32 float4s
    // valid_in_dis_idx == valid_idx for all 32 queries: computed once, stored twice
    bool same_l = true;
    for (int i = lane; i < nrows; i += 32) same_l = same_l && nv[i] == ns[i];
    const bool same = __all_sync(FULL_MASK, same_l);
    for (int fi = 0; fi < nfi; ++fi) {
        const int f = 32 * fi + lane;
        const bool on = f < kt;
        unsigned pos0 = 0, pos3 = 0;
        int r0 = 0, r3 = 0;
        if (on) {
            r0 = (int)udiv_magic(4u * f, magic_kt, (unsigned)kt, pos0);
            r3 = (int)udiv_magic(4u * f + 3u, magic_kt, (unsigned)kt, pos3);
        }
        // warp-uniform shortcut: no lane of the chunk straddles two rows (one term per element instead of two)
        const bool straddle = __any_sync(FULL_MASK, on && r0 != r3);
        if (!on) continue;
        // element i sits in row r0 at pos0 + i while that is < kt, else in row r3 at pos0 + i - kt: its count is read
        // from row r0 + sh[i] of the block and its position is pe[i] -- one shared-memory load and one FADD.SAT per
        // element when a lane of the chunk straddles two rows, one load per float4 when none does
        float pe[4];
        int sh[4];                                  // row of element i relative to r0 (0 or r3 - r0)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool lo = (int)pos0 + i < kt;
            pe[i] = (float)(lo ? (int)pos0 + i : (int)pos0 + i - kt);
            sh[i] = lo ? 0 : r3 - r0;
        }
        // the block of four rows advances every pointer by kt float4s = 16 kt bytes: a warp-uniform byte offset
        // added to the lane's fixed column address
        char* const ovb = reinterpret_cast<char*>(o_valid) + (size_t)f * 16;
        char* const odb = reinterpret_cast<char*>(o_vdis) + (size_t)f * 16;
        const unsigned step = (unsigned)kt * 16u;
        const float* nva = nv + r0;
        const float* nsa = ns + r0;
        auto body = [&](auto STR, auto SAME, auto BOTH) {
            constexpr bool two = decltype(STR)::value, one_array = decltype(SAME)::value, both = decltype(BOTH)::value;
#pragma unroll 8
            for (int blk = 0; blk < nblk; ++blk) {
                float a4[4], d4[4];
                if constexpr (two) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) a4[i] = __saturatef(nva[4 * blk + sh[i]] - pe[i]);
                } else {
                    const float va = nva[4 * blk];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a4[i] = __saturatef(va - pe[i]);
                }
                if constexpr (one_array) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) d4[i] = a4[i];
                } else if constexpr (two) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) d4[i] = __saturatef(nsa[4 * blk + sh[i]] - pe[i]);
                } else {
                    const float da = nsa[4 * blk];
#pragma unroll
                    for (int i = 0; i < 4; ++i) d4[i] = __saturatef(da - pe[i]);
                }
                const unsigned off = (unsigned)blk * step;
                if (both || o_valid) __stcs(reinterpret_cast<float4*>(ovb + off), make_float4(a4[0], a4[1], a4[2], a4[3]));
                if (both || o_vdis) __stcs(reinterpret_cast<float4*>(odb + off), make_float4(d4[0], d4[1], d4[2], d4[3]));
            }
        };
        auto pick = [&](auto STR, auto SAME) {
            if (o_valid && o_vdis) body(STR, SAME, std::true_type{}); else body(STR, SAME, std::false_type{});
        };
        if (straddle) { if (same) pick(std::true_type{}, std::true_type{}); else pick(std::true_type{}, std::false_type{}); }
        else          { if (same) pick(std::false_type{}, std::true_type{}); else pick(std::false_type{}, std::false_type{}); }
    }
    // rows that do not fill a block of four (or unaligned outputs): element by element
    const unsigned nel = (unsigned)nrows * kt;
    for (unsigned e = (unsigned)nblk * 4u * kt + lane; e < nel; e += 32) {
        unsigned pos;
        const unsigned r = udiv_magic(e, magic_kt, (unsigned)kt, pos);
        if (o_valid) o_valid[e] = (float)pos < nv[r] ? 1.0f : 0.0f;
        if (o_vdis) o_vdis[e] = (float)pos < ns[r] ? 1.0f : 0.0f;
    }
}

}  // namespace elo
