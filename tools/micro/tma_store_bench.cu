// Microbenchmark: throughput of TMA bulk stores (cp.async.bulk.global.shared::cta) as a function of copy size and of
// the relative alignment of source and destination, against plain 16-byte streaming stores.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_store_bench tma_store_bench.cu && ./tma_store_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// every thread issues `per_thread` copies of `bytes` from shared memory (offset src_off) to consecutive
// destinations; dst_off shifts the whole output (both multiples of 16)
__global__ void k_bulk(float* out, int bytes, int per_thread, int src_off, int dst_off, int stride_bytes)
{
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const unsigned s = (unsigned)__cvta_generic_to_shared(sm) + src_off;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned char* base = reinterpret_cast<unsigned char*>(out) + dst_off;
    for (int i = 0; i < per_thread; ++i) {
        const long long idx = t * per_thread + i;
        bulk_s2g(base + idx * stride_bytes, s + ((threadIdx.x * 16) & 1023), bytes);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// one elected thread per warp issues the copies for the warp
__global__ void k_bulk_warp(float* out, int bytes, int per_warp, int src_off, int dst_off)
{
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const unsigned s = (unsigned)__cvta_generic_to_shared(sm) + src_off;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned char* base = reinterpret_cast<unsigned char*>(out) + dst_off;
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < per_warp; ++i) bulk_s2g(base + (w * per_warp + i) * (long long)bytes, s, bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

__global__ void k_stg(float4* out, long long n4)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
        __stcs(out + i, make_float4(1.f, 1.f, 1.f, 1.f));
}

template <class F>
static float time_it(F f, int iters = 10)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / iters;
}

int main()
{
    const long long total = 161280000ll;          // bytes of the two count tensors of configs[0]
    float* out;
    cudaMalloc(&out, total + (1 << 20));
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
    cudaFuncSetAttribute(k_bulk_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    {
        const float us = time_it([&] { k_stg<<<148 * 8, 256>>>(reinterpret_cast<float4*>(out), total / 16); });
        printf("STG.128 .cs grid-stride            : %8.1f us  %7.1f GB/s\n", us, total / us / 1e3);
    }
    // per-thread copies, 720 CTAs x 160 threads = 115200 threads
    const int sizes[] = {688, 1376, 2752};
    for (int bytes : sizes) {
        const int per_thread = (int)(total / 115200 / bytes);
        for (int src_off : {0, 16, 64})
            for (int dst_off : {0, 16}) {
                const float us = time_it([&] { k_bulk<<<720, 160, 20000>>>(out, bytes, per_thread, src_off, dst_off, bytes); });
                const double mb = 115200.0 * per_thread * bytes;
                printf("bulk per-thread %5d B x %d  src+%-3d dst+%-3d: %8.1f us  %7.1f GB/s\n", bytes, per_thread, src_off, dst_off, us, mb / us / 1e3);
            }
    }
    // 700-byte rows (unaligned size is illegal) -> 704-byte stride, 688-byte copies: as the kernel does
    {
        const float us = time_it([&] { k_bulk<<<720, 160, 20000>>>(out, 688, 2, 0, 0, 704); });
        printf("bulk per-thread 688 B x 2, stride 704          : %8.1f us\n", us);
    }
    // per-warp elected copies, larger blocks
    for (int bytes : {2816, 5632, 11264, 22528, 45056}) {
        const long long warps = 720ll * 5;
        const int per_warp = (int)(total / warps / bytes);
        for (int src_off : {0, 16})
            for (int dst_off : {0, 16}) {
                const float us = time_it([&] { k_bulk_warp<<<720, 160, 70000>>>(out, bytes, per_warp, src_off, dst_off); });
                const double mb = (double)warps * per_warp * bytes;
                printf("bulk per-warp   %5d B x %d src+%-3d dst+%-3d: %8.1f us  %7.1f GB/s\n", bytes, per_warp, src_off, dst_off, us, mb / us / 1e3);
            }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
