// project_pose.cu -- the per-point parts of the pose warp-refinement (sm_100a):
//
//   elo_project     PreProcess crop/augment (model_util.py:346-445) or quaternion warp
//                   (model_util.py:17-69, pwclo_model.py:213-227) fused with ProjectPC2SphericalRing
//                   (model_util.py:181-292): spherical binning, min-range winner per cell, scatter.
//   elo_pose_head   softmax_valid (model_util.py:319-343) + the conv1d pose heads and the pose
//                   composition (pwclo_model.py:194-208, 262-280).
//
// The reference builds these from ~30 elementwise TensorFlow kernels per level plus unique /
// unsorted_segment_min / scatter_nd inside a Python loop over the batch.  TensorFlow evaluates every
// multiply and add as its own kernel, so nothing is FMA-contracted there; the arithmetic that decides
// a bin or a range winner is therefore written with explicit round-to-nearest intrinsics in the
// reference's operation order, which makes the integer outcomes (cells, winners) reproducible.
#include <cuda_runtime.h>

#include <algorithm>
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/elo_b200.h"
#include "elo_common.cuh"

namespace elo {

constexpr int PROJECT_CTAS_THROUGHPUT = 2;    // CTAs per SM of the projection passes under the throughput tile policy

struct ProjParams {
    int B, N, H, W, C, mode;
    const float* points; long long point_stride, batch_stride, outer_stride; int inner_batch;
    const float* feat;
    const float* T; const int* T_apply; const float* q; const float* t;
    float pi, az, vres, voff;
    unsigned long long* cellmin;   // (B,H,W): (~epoch << 32) | range bits of the nearest point seen in this epoch
    unsigned* state;               // [0] epoch of the current call, [1] CTAs of the scatter kernel that are done
    float* out_xyz; float* out_feat; float* out_points;
    int* out_cell;       // optional (B, N): cell of the point if it is (one of) the nearest of its cell, else -1
    int2* keys;          // optional (B, N): (cell, range bits) of every point, bin kernel -> scatter kernel
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

// Hamilton product a (x) b, term order as written in model_util.py:21-31
__device__ __forceinline__ void hamilton(const float a[4], const float b[4], float r[4])
{
    r[0] = sub(sub(sub(mul(a[0], b[0]), mul(a[1], b[1])), mul(a[2], b[2])), mul(a[3], b[3]));
    r[1] = sub(add(add(mul(a[0], b[1]), mul(a[1], b[0])), mul(a[2], b[3])), mul(a[3], b[2]));
    r[2] = add(add(sub(mul(a[0], b[2]), mul(a[1], b[3])), mul(a[2], b[0])), mul(a[3], b[1]));
    r[3] = add(sub(add(mul(a[0], b[3]), mul(a[1], b[2])), mul(a[2], b[1])), mul(a[3], b[0]));
}

// conj(q) / (|q|^2 + 1e-10), model_util.py:61-69
__device__ __forceinline__ void inv_q(const float q[4], float r[4])
{
    const float n2 = add(add(add(add(mul(q[0], q[0]), mul(q[1], q[1])), mul(q[2], q[2])), mul(q[3], q[3])), 1e-10f);
    r[0] = __fdiv_rn(q[0], n2); r[1] = __fdiv_rn(-q[1], n2); r[2] = __fdiv_rn(-q[2], n2); r[3] = __fdiv_rn(-q[3], n2);
}

// The point as the network sees it after PreProcess (mode 1) or after the pose warp (mode 2).
__device__ __forceinline__ void transform_point(const ProjParams& p, int b, int n, float& x, float& y, float& z)
{
    const float* src = p.points + (size_t)n * p.point_stride +
                       (p.inner_batch > 0 ? (size_t)(b % p.inner_batch) * p.batch_stride + (size_t)(b / p.inner_batch) * p.outer_stride
                                          : (size_t)b * p.batch_stride);
    x = __ldg(src); y = __ldg(src + 1); z = __ldg(src + 2);
    const bool valid = !(x == 0.f && y == 0.f && z == 0.f);
    if (p.mode == 1) {
        // 35 m crop on the xy range, then the augmentation matrix on [p, 1] for the frame that is
        // augmented (the other frame is NOT multiplied, not even by an identity), then * valid
        // (model_util.py:380-383, 390-417, 419-420).  The final multiply keeps the sign of a zero
        // coordinate (-0.0 * 0 = -0.0), which atan2 turns into azimuth pi instead of 0: kept as is.
        const float rxy = __fsqrt_rn(add(mul(x, x), mul(y, y)));
        float w = 1.f;
        if (rxy > 35.f) { x = 0.f; y = 0.f; z = 0.f; w = 0.f; }
        if (p.T != nullptr && (p.T_apply == nullptr || __ldg(p.T_apply + b) != 0)) {
            const float* T = p.T + (size_t)b * 16;
            const float nx = add(add(add(mul(T[0], x), mul(T[1], y)), mul(T[2], z)), mul(T[3], w));
            const float ny = add(add(add(mul(T[4], x), mul(T[5], y)), mul(T[6], z)), mul(T[7], w));
            const float nz = add(add(add(mul(T[8], x), mul(T[9], y)), mul(T[10], z)), mul(T[11], w));
            x = nx; y = ny; z = nz;
        }
        const float m = valid ? 1.f : 0.f;
        x = mul(x, m); y = mul(y, m); z = mul(z, m);
    } else if (p.mode == 2) {
        float q[4], qi[4], pq[4] = {0.f, x, y, z}, a[4], r[4];
        for (int i = 0; i < 4; ++i) q[i] = __ldg(p.q + (size_t)b * 4 + i);
        inv_q(q, qi);
        hamilton(q, pq, a);
        hamilton(a, qi, r);
        const float m = valid ? 1.f : 0.f;
        x = mul(add(r[1], __ldg(p.t + (size_t)b * 3 + 0)), m);
        y = mul(add(r[2], __ldg(p.t + (size_t)b * 3 + 1)), m);
        z = mul(add(r[3], __ldg(p.t + (size_t)b * 3 + 2)), m);
    }
}

// cell and range of a point (model_util.py:225-245).  float->int casts truncate; NaN casts to 0 on the
// GPU, which is what sends the zero points (asin(0/0)) to row H-1 in the reference graph.
__device__ __forceinline__ int bin_point(const ProjParams& p, float x, float y, float z, float& r)
{
    r = __fsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z)));
    int col = __float2int_rz(__fdiv_rn(sub(p.pi, atan2f(y, x)), p.az));
    const float beta = asinf(__fdiv_rn(z, r));
    const int tmp = __float2int_rz(add(__fdiv_rn(beta, p.vres), p.voff));
    int row = p.H - tmp;
    row = min(max(row, 0), p.H - 1);
    col = min(max(col, 0), p.W - 1);
    return row * p.W + col;
}

// Cell minima carry the call's epoch in their high word ((~epoch) << 32 | range bits): a value left by an
// earlier call is larger than anything this call writes, so the table never has to be cleared -- the separate
// initialisation launch is gone (the binning kernel also zeroes the images the scatter adds into).  The epoch
// lives in device memory (state[0]) and is advanced by the last CTA of the scatter kernel, which keeps a captured
// CUDA graph replayable.
__device__ __forceinline__ unsigned long long cell_key(unsigned epoch, unsigned rbits)
{
    return ((unsigned long long)(~epoch) << 32) | rbits;
}

__global__ void project_bin_kernel(const ProjParams p)
{
    pdl_trigger();
    pdl_wait();          // inputs come from earlier kernels; the output images may alias memory they still read
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(p.state);
    const long long total = (long long)p.B * p.N;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    {
        const long long cells = (long long)p.B * p.H * p.W;
        for (long long c = i0; c < cells * 3; c += stride) p.out_xyz[c] = 0.f;
        if (p.out_feat != nullptr)
            for (long long c = i0; c < cells * p.C; c += stride) p.out_feat[c] = 0.f;
    }
    // A third of a frame's points are zero padding or cropped away: exactly (+0, +0, +0) after the transform.  Their
    // cell is a constant of the image geometry; one thread per CTA evaluates it with the very arithmetic the other
    // points go through, and the padding skips atan2 / asin / sqrt / the divisions.
    __shared__ int s_zero_cell;
    if (threadIdx.x == 0) {
        float r0;
        s_zero_cell = bin_point(p, 0.f, 0.f, 0.f, r0);
    }
    __syncthreads();
    const int zero_cell = s_zero_cell;
    for (long long i = i0; i < total; i += stride) {
        const int b = (int)(i / p.N), n = (int)(i % p.N);
        float x, y, z, r;
        transform_point(p, b, n, x, y, z);
        int cell;
        if ((__float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z)) == 0u) { cell = zero_cell; r = 0.f; }
        else cell = bin_point(p, x, y, z, r);
        // r >= 0 (or NaN, which orders above every finite value as an unsigned pattern).  Zero-padded /
        // cropped points all fall into ONE cell per sample with r = 0, the smallest possible key: a plain
        // store is enough for them and avoids ~10^5 serialised atomics on a single address.
        unsigned long long* slot = p.cellmin + (size_t)b * p.H * p.W + cell;
        const unsigned rbits = __float_as_uint(r);
        if (p.keys != nullptr) p.keys[i] = make_int2(cell, (int)rbits);
        if (rbits == 0u) *slot = cell_key(epoch, 0u);
        else atomicMin(slot, cell_key(epoch, rbits));
        if (p.out_points != nullptr) {
            float* o = p.out_points + (size_t)i * 3;
            o[0] = x; o[1] = y; o[2] = z;
        }
    }
}

// one thread per (point, 4-channel slab): slab 0 carries xyz, slabs 1.. the features
__global__ void project_scatter_kernel(const ProjParams p)
{
    pdl_trigger();
    pdl_wait();
    const unsigned epoch = *reinterpret_cast<volatile unsigned*>(p.state);
    const int slabs = 1 + (p.out_feat != nullptr ? (p.C + 3) / 4 : 0);
    const long long total = (long long)p.B * p.N * slabs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long pt = i / slabs;
        const int slab = (int)(i - pt * slabs);
        const int b = (int)(pt / p.N), n = (int)(pt % p.N);
        float x = 0.f, y = 0.f, z = 0.f, r;
        int cell;
        unsigned rbits;
        if (p.keys != nullptr) {                    // left by the binning pass: no second transform / binning
            const int2 k = p.keys[pt];
            cell = k.x; rbits = (unsigned)k.y;
        } else {
            transform_point(p, b, n, x, y, z);
            cell = bin_point(p, x, y, z, r);
            rbits = __float_as_uint(r);
        }
        const size_t gcell = (size_t)b * p.H * p.W + cell;
        const bool winner = cell_key(epoch, rbits) == p.cellmin[gcell];
        if (slab == 0 && p.out_cell != nullptr) p.out_cell[pt] = winner ? cell : -1;
        if (!winner) continue;                                       // not the (a) nearest point of its cell
        if (slab == 0) {
            if (p.keys != nullptr) transform_point(p, b, n, x, y, z);     // winners only: one point per cell
            // equal-range ties accumulate, like scatter_nd (model_util.py:271)
            if (x != 0.f) atomicAdd(p.out_xyz + gcell * 3 + 0, x);
            if (y != 0.f) atomicAdd(p.out_xyz + gcell * 3 + 1, y);
            if (z != 0.f) atomicAdd(p.out_xyz + gcell * 3 + 2, z);
        } else {
            const int c0 = (slab - 1) * 4;
            const float* f = p.feat + ((size_t)b * p.N + n) * p.C;
            for (int c = c0; c < min(c0 + 4, p.C); ++c) {
                const float v = __ldg(f + c);
                if (v != 0.f) atomicAdd(p.out_feat + gcell * p.C + c, v);
            }
        }
    }
    // every CTA has read the epoch by now: the last one to finish advances it for the next call on this table
    __shared__ bool s_last_cta;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last_cta = atomicAdd(p.state + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last_cta) return;
    // The 32-bit epoch wraps after 2^32 calls on one table; keys of epoch 0 (high word ~0) would then lose against
    // every stale minimum.  On the wrapping call the last CTA -- every other CTA is done with the table -- refills
    // it with all ones, the state a fresh table starts from.
    if (epoch + 1u == 0u) {
        const long long cells = (long long)p.B * p.H * p.W;
        for (long long c = threadIdx.x; c < cells; c += blockDim.x) p.cellmin[c] = ~0ull;
        __threadfence();
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.state[1] = 0u;
        p.state[0] = epoch + 1u;
        __threadfence();
    }
}

// ================================================================================================
struct PoseParams {
    int B, N, G, has_coarse;
    const float* feature;   // (B, N, 64)
    const float* weight;    // (B, N, 64) attention logits (the embedding mask)
    const float* xyz;       // (B, N, 3): a point is valid unless it is exactly (0,0,0)
    const float* w_big; const float* b_big;   // (64, 256), (256)
    const float* w_q; const float* b_q;       // (256, 4), (4)
    const float* w_t; const float* b_t;       // (256, 3), (3)
    const float* q_coarse; const float* t_coarse;   // (B,4), (B,3)
    float* partial;         // (B, G, 3, 64) scratch
    unsigned* counter;      // (B) zero before the first launch; left zero afterwards
    float* q_out; float* t_out; float* q_norm_out; float* pooled_out;
};

constexpr int POSE_THREADS = 256;
constexpr int POSE_PT = 8;          // points per thread; a CTA covers 4 * POSE_PT = 32 points

__global__ void __launch_bounds__(POSE_THREADS) pose_head_kernel(const PoseParams p)
{
    pdl_trigger();
    __shared__ float s_m[4][64], s_s[4][64], s_a[4][64];
    __shared__ float s_pool[64], s_big[256], s_head[8];
    __shared__ bool s_last;
    const int b = blockIdx.y, gidx = blockIdx.x;
    const int c = threadIdx.x & 63, sub = threadIdx.x >> 6;
    // Head weights are constants: every CTA starts fetching its threads' share now (whichever CTA turns out to
    // be the last one of the sample then runs the heads from registers, with no dependent global loads left).
    float wbig[64], whead[8];
    float bbig = 0.f, bhead = 0.f;
    if (p.w_big != nullptr) {
        const int o = threadIdx.x;
#pragma unroll
        for (int k = 0; k < 64; ++k) wbig[k] = __ldg(p.w_big + k * 256 + o);
        bbig = __ldg(p.b_big + o);
        const int hw_ = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            whead[j] = hw_ < 4 ? __ldg(p.w_q + (ln + 32 * j) * 4 + hw_) : (hw_ < 7 ? __ldg(p.w_t + (ln + 32 * j) * 3 + (hw_ - 4)) : 0.f);
        bhead = hw_ < 4 ? __ldg(p.b_q + hw_) : (hw_ < 7 ? __ldg(p.b_t + (hw_ - 4)) : 0.f);
    }
    pdl_wait();          // features / logits / coarse pose come from the kernels before

    // online masked softmax over this CTA's slice of the points, per channel; the slice is short
    // (<= 4 * POSE_PT points) and all of a thread's loads are issued before any of them is used
    const int per = (p.N + p.G - 1) / p.G;
    const int n0 = gidx * per, n1 = min(p.N, n0 + per);
    float wv[POSE_PT], fv[POSE_PT];
    bool ok[POSE_PT];
#pragma unroll
    for (int i = 0; i < POSE_PT; ++i) {
        const int n = n0 + sub + 4 * i;
        ok[i] = n < n1;
        wv[i] = 0.f; fv[i] = 0.f;
        if (ok[i]) {
            const float* xyz = p.xyz + ((size_t)b * p.N + n) * 3;
            ok[i] = !(__ldg(xyz) == 0.f && __ldg(xyz + 1) == 0.f && __ldg(xyz + 2) == 0.f);
            wv[i] = __ldg(p.weight + ((size_t)b * p.N + n) * 64 + c);
            fv[i] = __ldg(p.feature + ((size_t)b * p.N + n) * 64 + c);
        }
    }
    float m = -INFINITY, s = 0.f, a = 0.f;
#pragma unroll
    for (int i = 0; i < POSE_PT; ++i)
        if (ok[i]) m = fmaxf(m, wv[i]);
#pragma unroll
    for (int i = 0; i < POSE_PT; ++i)
        if (ok[i]) {
            const float e = expf(wv[i] - m);
            s += e;
            a = fmaf(e, fv[i], a);
        }
    s_m[sub][c] = m; s_s[sub][c] = s; s_a[sub][c] = a;
    __syncthreads();
    if (sub == 0) {
        float M = fmaxf(fmaxf(s_m[0][c], s_m[1][c]), fmaxf(s_m[2][c], s_m[3][c]));
        float S = 0.f, A = 0.f;
        for (int i = 0; i < 4; ++i) {
            const float sc = s_m[i][c] == -INFINITY ? 0.f : expf(s_m[i][c] - M);
            S = fmaf(s_s[i][c], sc, S);
            A = fmaf(s_a[i][c], sc, A);
        }
        float* part = p.partial + ((size_t)b * p.G + gidx) * 192;
        part[c] = M; part[64 + c] = S; part[128 + c] = A;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(p.counter + b, 1u) == (unsigned)(p.G - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // last CTA of the sample: combine the partials (all 256 threads: 4 strided groups per channel,
    // loads issued together), then run the heads
    {
        const float* part = p.partial + (size_t)b * p.G * 192;
        float M = -INFINITY, S = 0.f, A = 0.f;
        for (int g0 = sub; g0 < p.G; g0 += 4 * 8) {
            float mg[8], sg[8], ag[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int g = g0 + 4 * u;
                const bool in = g < p.G;
                mg[u] = in ? __ldcg(part + g * 192 + c) : -INFINITY;
                sg[u] = in ? __ldcg(part + g * 192 + 64 + c) : 0.f;
                ag[u] = in ? __ldcg(part + g * 192 + 128 + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (mg[u] == -INFINITY) continue;
                const float Mn = fmaxf(M, mg[u]);
                const float s0 = M == -INFINITY ? 0.f : expf(M - Mn), s1 = expf(mg[u] - Mn);
                S = S * s0 + sg[u] * s1;
                A = A * s0 + ag[u] * s1;
                M = Mn;
            }
        }
        __syncthreads();
        s_m[sub][c] = M; s_s[sub][c] = S; s_a[sub][c] = A;
        __syncthreads();
    }
    if (threadIdx.x < 64) {
        float M = fmaxf(fmaxf(s_m[0][c], s_m[1][c]), fmaxf(s_m[2][c], s_m[3][c]));
        float S = 0.f, A = 0.f;
        for (int i = 0; i < 4; ++i) {
            const float sc = s_m[i][c] == -INFINITY ? 0.f : expf(s_m[i][c] - M);
            S = fmaf(s_s[i][c], sc, S);
            A = fmaf(s_a[i][c], sc, A);
        }
        const float v = A / S;      // no valid point at all: 0/0 = NaN, as an empty softmax gives upstream
        s_pool[c] = v;
        if (p.pooled_out != nullptr) p.pooled_out[(size_t)b * 64 + c] = v;
    }
    if (threadIdx.x == 0) p.counter[b] = 0u;
    __syncthreads();
    if (p.w_big == nullptr) return;     // softmax_valid only
    {   // conv1d 64 -> 256, no activation (pwclo_model.py:197); dropout is the identity at inference
        float acc = bbig;
#pragma unroll
        for (int k = 0; k < 64; ++k) acc = fmaf(s_pool[k], wbig[k], acc);
        s_big[threadIdx.x] = acc;
    }
    __syncthreads();
    {   // 7 heads of 256 MACs: warp w computes head w (q0..q3, t0..t2)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp < 7) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(s_big[lane + 32 * j], whead[j], acc);
            for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_head[warp] = acc + bhead;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float q[4] = {s_head[0], s_head[1], s_head[2], s_head[3]};
        float t[3] = {s_head[4], s_head[5], s_head[6]};
        // q / (sqrt(sum q^2 + 1e-10) + 1e-10)   (pwclo_model.py:203)
        const float nq = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + 1e-10f) + 1e-10f;
        for (int i = 0; i < 4; ++i) q[i] = q[i] / nq;
        float qo[4], to[3];
        if (p.has_coarse) {
            // q = q_det (x) q_coarse ;  t = (q_det (x) [0,t_coarse] (x) q_det^-1)[1:] + t_det   (:273-280)
            float qc[4], tc[4] = {0.f, 0.f, 0.f, 0.f}, qi[4], a4[4], r4[4];
            for (int i = 0; i < 4; ++i) qc[i] = p.q_coarse[(size_t)b * 4 + i];
            for (int i = 0; i < 3; ++i) tc[i + 1] = p.t_coarse[(size_t)b * 3 + i];
            hamilton(q, tc, a4);
            inv_q(q, qi);
            hamilton(a4, qi, r4);
            hamilton(q, qc, qo);
            for (int i = 0; i < 3; ++i) to[i] = r4[i + 1] + t[i];
        } else {
            for (int i = 0; i < 4; ++i) qo[i] = q[i];
            for (int i = 0; i < 3; ++i) to[i] = t[i];
        }
        const float no = sqrtf(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3] + 1e-10f) + 1e-10f;
        for (int i = 0; i < 4; ++i) {
            p.q_out[(size_t)b * 4 + i] = qo[i];
            p.q_norm_out[(size_t)b * 4 + i] = qo[i] / no;
        }
        for (int i = 0; i < 3; ++i) p.t_out[(size_t)b * 3 + i] = to[i];
    }
}

// Strided xyz pyramid (pwclo_model.py:88-114): level l keeps pixel (i*sh_l, j*sw_l) of the input image.
struct PyramidParams {
    int S, H, W;
    int oh[4], ow[4], sh[4], sw[4];
    long long start[5];          // first cell of each level in the flat work list (per sample)
    const float* in;
    float* out[4];
};

__global__ void pyramid_xyz_kernel(const PyramidParams p)
{
    pdl_trigger();
    pdl_wait();
    const long long per = p.start[4];
    const long long total = per * p.S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(i / per);
        const long long r = i - (long long)s * per;
        int l = 0;
        while (l < 3 && r >= p.start[l + 1]) ++l;
        const int cell = (int)(r - p.start[l]);
        const int y = cell / p.ow[l], x = cell - y * p.ow[l];
        const float* src = p.in + (((size_t)s * p.H + (size_t)y * p.sh[l]) * p.W + (size_t)x * p.sw[l]) * 3;
        float* dst = p.out[l] + ((size_t)s * p.oh[l] * p.ow[l] + cell) * 3;
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2);
    }
}

__global__ void gt_pose_kernel(int B, const float* T_gt, const float* T_trans, const float* T_trans_inv,
                               const int* aug_frame, float* q_gt, float* t_gt)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int mode = aug_frame ? aug_frame[b] : 2;
    const float* A = mode == 2 ? T_trans + b * 16 : T_gt + b * 16;
    const float* Bm = mode == 2 ? T_gt + b * 16 : T_trans_inv + b * 16;
    float T[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) acc = add(acc, mul(A[i * 4 + k], Bm[k * 4 + j]));
            T[i * 4 + j] = acc;
        }
    // mat2euler (zyx) then euler2quat, model_util.py:111-142
    const float cy = sqrtf(T[10] * T[10] + T[6] * T[6]);
    const float z = atan2f(-T[1], T[0]) * 0.5f, y = atan2f(T[2], cy) * 0.5f, x = atan2f(-T[6], T[10]) * 0.5f;
    const float cz = cosf(z), sz = sinf(z), cyy = cosf(y), sy = sinf(y), cx = cosf(x), sx = sinf(x);
    q_gt[b * 4 + 0] = cx * cyy * cz - sx * sy * sz;
    q_gt[b * 4 + 1] = cx * sy * sz + cyy * cz * sx;
    q_gt[b * 4 + 2] = cx * cz * sy - sx * cyy * sz;
    q_gt[b * 4 + 3] = cx * cyy * sz + sx * cz * sy;
    t_gt[b * 3 + 0] = T[3]; t_gt[b * 3 + 1] = T[7]; t_gt[b * 3 + 2] = T[11];
}

}  // namespace elo

using namespace elo;

extern "C" int elo_pyramid_xyz(int samples, int H, int W, const int* out_h, const int* out_w, const int* stride_h,
                               const int* stride_w, const float* xyz_in, float* const* out, void* stream)
{
    if (samples < 0 || H <= 0 || W <= 0 || !out_h || !out_w || !stride_h || !stride_w || !xyz_in || !out)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "pyramid_xyz: bad arguments");
    if (samples == 0) return ELO_OK;
    PyramidParams p;
    p.S = samples; p.H = H; p.W = W; p.in = xyz_in;
    p.start[0] = 0;
    for (int l = 0; l < 4; ++l) {
        if (out_h[l] <= 0 || out_w[l] <= 0 || stride_h[l] <= 0 || stride_w[l] <= 0 || !out[l] ||
            (long long)(out_h[l] - 1) * stride_h[l] >= H || (long long)(out_w[l] - 1) * stride_w[l] >= W)
            return set_error(ELO_ERR_INVALID_ARGUMENT, "pyramid_xyz: level outside the input image");
        p.oh[l] = out_h[l]; p.ow[l] = out_w[l]; p.sh[l] = stride_h[l]; p.sw[l] = stride_w[l]; p.out[l] = out[l];
        p.start[l + 1] = p.start[l] + (long long)out_h[l] * out_w[l];
    }
    const long long total = p.start[4] * samples;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    cudaError_t err = launch(pyramid_xyz_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, p);
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "pyramid_xyz launch");
}

extern "C" int elo_gt_pose(int batch_size, const float* T_gt, const float* T_trans, const float* T_trans_inv,
                           const int* aug_frame, float* q_gt, float* t_gt, void* stream)
{
    if (batch_size < 0 || !T_gt || !T_trans || !T_trans_inv || !q_gt || !t_gt)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "gt_pose: bad arguments");
    if (batch_size == 0) return ELO_OK;
    gt_pose_kernel<<<(batch_size + 63) / 64, 64, 0, (cudaStream_t)stream>>>(batch_size, T_gt, T_trans, T_trans_inv,
                                                                           aug_frame, q_gt, t_gt);
    count_launches(1);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "gt_pose launch");
}

extern "C" int elo_project(const elo_project_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "project: null descriptor");
    if (d->batch_size < 0 || d->num_points <= 0 || d->H <= 1 || d->W <= 0 || !d->points || !d->cellmin || !d->state || !d->out_xyz ||
        d->mode < 0 || d->mode > 2 || d->point_stride < 3 || (d->feat != nullptr && (d->C <= 0 || !d->out_feat)) ||
        (d->mode == 2 && (!d->q || !d->t)))
        return set_error(ELO_ERR_INVALID_ARGUMENT, "project: bad arguments");
    if (d->batch_size == 0) return ELO_OK;
    ProjParams p;
    p.B = d->batch_size; p.N = d->num_points; p.H = d->H; p.W = d->W; p.C = d->feat ? d->C : 0; p.mode = d->mode;
    p.points = d->points; p.point_stride = d->point_stride; p.batch_stride = d->batch_stride;
    p.inner_batch = d->inner_batch; p.outer_stride = d->outer_stride;
    p.feat = d->feat; p.T = d->T; p.T_apply = d->T_apply; p.q = d->q; p.t = d->t;
    p.pi = d->pi; p.az = d->az_res; p.vres = d->v_res; p.voff = d->v_off;
    p.cellmin = d->cellmin; p.state = d->state; p.out_xyz = d->out_xyz; p.out_feat = d->feat ? d->out_feat : nullptr;
    p.out_points = d->out_points;
    p.out_cell = d->out_cell;
    p.keys = reinterpret_cast<int2*>(d->point_keys);
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = device_info().sm_count;
    // CTAs per SM: 8 x 256 threads fill an SM's thread slots -- the shortest kernel, but both passes wait on memory
    // (atomics, scattered loads) most of the time.  Under the throughput tile policy fewer CTAs with more points per
    // thread leave the slots to the other forwards in flight (ELO_PROJECT_CAP overrides).
    static const int cap_env = getenv("ELO_PROJECT_CAP") ? atoi(getenv("ELO_PROJECT_CAP")) : 0;
    const int per_sm = cap_env > 0 ? cap_env : (elo_get_tile_policy() == 1 ? PROJECT_CTAS_THROUGHPUT : 8);
    auto blocks = [&](long long work) {
        long long b = (work + 255) / 256;
        const long long cap = (long long)sms * per_sm;
        return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    const long long cells = (long long)p.B * p.H * p.W;
    const long long bin_work = std::max((long long)p.B * p.N, cells * (p.C > 3 ? p.C : 3));   // points, and image zeroing
    cudaError_t err = launch(project_bin_kernel, dim3(blocks(bin_work)), dim3(256), 0, st, p);
    const int slabs = 1 + (p.out_feat ? (p.C + 3) / 4 : 0);
    if (err == cudaSuccess)
        err = launch(project_scatter_kernel, dim3(blocks((long long)p.B * p.N * slabs)), dim3(256), 0, st, p);
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "project launch");
}

extern "C" int elo_pose_head(const elo_pose_head_desc* d, void* stream)
{
    if (d == nullptr) return set_error(ELO_ERR_INVALID_ARGUMENT, "pose_head: null descriptor");
    const bool heads = d->w_big != nullptr;
    if (d->batch_size < 0 || d->num_points <= 0 || !d->feature || !d->weight || !d->xyz || !d->partial ||
        !d->counter || d->num_slices < 1 || (!heads && !d->pooled_out) ||
        (heads && (!d->b_big || !d->w_q || !d->b_q || !d->w_t || !d->b_t || !d->q_out || !d->t_out ||
                   !d->q_norm_out || (d->has_coarse && (!d->q_coarse || !d->t_coarse)))))
        return set_error(ELO_ERR_INVALID_ARGUMENT, "pose_head: bad arguments");
    if (d->batch_size == 0) return ELO_OK;
    if ((long long)d->num_slices * 4 * POSE_PT < d->num_points)
        return set_error(ELO_ERR_INVALID_ARGUMENT, "pose_head: num_slices must be at least ceil(num_points / 32)");
    PoseParams p;
    p.B = d->batch_size; p.N = d->num_points; p.G = d->num_slices; p.has_coarse = d->has_coarse;
    p.feature = d->feature; p.weight = d->weight; p.xyz = d->xyz;
    p.w_big = d->w_big; p.b_big = d->b_big; p.w_q = d->w_q; p.b_q = d->b_q; p.w_t = d->w_t; p.b_t = d->b_t;
    p.q_coarse = d->q_coarse; p.t_coarse = d->t_coarse; p.partial = d->partial; p.counter = d->counter;
    p.q_out = d->q_out; p.t_out = d->t_out; p.q_norm_out = d->q_norm_out; p.pooled_out = d->pooled_out;
    dim3 grid(p.G, p.B);
    cudaError_t err = launch(pose_head_kernel, grid, dim3(POSE_THREADS), 0, (cudaStream_t)stream, p);
    return err == cudaSuccess ? ELO_OK : set_cuda_error(err, "pose_head launch");
}
