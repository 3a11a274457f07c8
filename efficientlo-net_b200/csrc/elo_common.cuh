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

}  // namespace elo
