// elo_common.cuh -- host-side plumbing shared by every translation unit of libelo_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace elo {

struct DeviceInfo {
    int device;
    int sm_count;          // 148 on B200
    int max_smem_optin;    // bytes of dynamic shared memory a CTA may opt in to (227 KB on sm_100)
};

// Properties of the current device (cached per device ordinal; no synchronisation).
const DeviceInfo& device_info();

// Record a message for elo_last_error() and return `code` (so callers can `return set_error(...)`).
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t err, const char* where);

// Every kernel launch of this library is counted (elo_launch_count() in the C ABI): bench.py reports it.
void count_launches(int n);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// A forward is a chain of ~40 dependent kernels, most of them a single wave of CTAs.  Every kernel of
// this library therefore (1) calls pdl_trigger() first thing, which lets the NEXT kernel in the stream be
// scheduled as soon as all of this kernel's CTAs are running, (2) does the part of its prologue that
// touches no upstream data (mbarrier init, TMEM allocation, weight / bias prefetch), and (3) calls
// pdl_wait() before the first access to anything an earlier kernel may have written (or may still be
// reading: scratch buffers are re-used).  pdl_wait() returns once every preceding kernel has completed
// and its writes are visible, so everything after it has plain stream-order semantics.
// Without the launch attribute (elo_set_pdl(0)) both are no-ops and launches serialise as usual.
bool pdl_enabled();

template <class Params, class Kernel>
cudaError_t launch(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Params& params)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    count_launches(1);
    return cudaLaunchKernelEx(&cfg, kernel, params);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

}  // namespace elo
