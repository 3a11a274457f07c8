// elo_tc.cuh -- 5th-generation tensor-core (tcgen05) primitives for the per-group MLPs, sm_100a.
//
// One CTA owns a 128-row tile.  A dense layer  D[128 x N] = A[128 x K] * W[K x N]  runs as
// tcgen05.mma.kind::tf32 with
//   * the accumulator D in tensor memory (TMEM): lane = row, column = output channel, fp32;
//   * the A operand ALSO in TMEM (the "TS" form): lane = row, column = input channel.  Activations
//     therefore never touch shared memory: an epilogue reads D with tcgen05.ld, applies bias + ReLU in
//     registers and writes the next layer's A operand back with tcgen05.st;
//   * the B operand (weights) in shared memory in the canonical K-major, no-swizzle layout
//     [K/4][N][4 floats]: a core matrix is 8 rows x 16 B, rows 16 B apart, 8-row groups 128 B apart
//     (stride byte offset), the two 16 B K-chunks of one K=8 MMA N*16 B apart (leading byte offset).
//
// fp32 accuracy from tf32 tensor cores: every operand is split x = hi + lo with hi = x truncated to
// tf32 (10 explicit mantissa bits) and lo = x - hi (exact in fp32), and a layer issues three MMAs per
// K-step into the same accumulator: A_hi*W_hi + A_lo*W_hi + A_hi*W_lo.  The dropped lo*lo term is
// ~2^-22 relative, which keeps the features inside the 1e-4 parity bar.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elo {
namespace tc {

constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a fully converged warp (elect.sync): lets ptxas treat what follows as warp-uniform, so
// tcgen05.mma is issued without a per-instruction ELECT / branch sequence
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- TMEM allocation (one warp, all lanes) -------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_holder)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, Blackwell version bit set
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: tf32 x tf32 -> f32, A and B K-major, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_tf32(uint32_t n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]; one thread issues
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// same with a compile-time accumulate flag: no predicate arithmetic in the issue loop
template <bool ACC>
__device__ __forceinline__ void mma_ts_c(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc)
{
    if (ACC)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem),
                     "l"(b_desc), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem),
                     "l"(b_desc), "r"(idesc) : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}

// ---- TMEM <-> registers (warp-collective; the warp's lane quarter is 32 * (warp % 4)) ------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x = hi + lo with hi carrying the tf32 part
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

}  // namespace tc
}  // namespace elo
