// elo_common.cu -- error reporting and device-property cache for libelo_b200.so.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/elo_b200.h"
#include "elo_common.cuh"

namespace elo {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_pdl{-1};     // -1: not decided yet (environment ELO_PDL, default on)
bool pdl_enabled()
{
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("ELO_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

int set_error(int code, const char* msg)
{
    snprintf(g_error, sizeof(g_error), "%s", msg);
    return code;
}

int set_cuda_error(cudaError_t err, const char* where)
{
    snprintf(g_error, sizeof(g_error), "%s: %s (%s)", where, cudaGetErrorName(err), cudaGetErrorString(err));
    return (int)err;
}

const DeviceInfo& device_info()
{
    static DeviceInfo cache[64];
    static bool have[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!have[dev]) {
        DeviceInfo d;
        d.device = dev;
        d.sm_count = 148;
        d.max_smem_optin = 227 * 1024;
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cache[dev] = d;
        have[dev] = true;
    }
    return cache[dev];
}

}  // namespace elo

extern "C" const char* elo_last_error(void) { return elo::g_error; }
extern "C" int elo_version(void) { return 100; }
extern "C" int elo_set_pdl(int on)
{
    elo::g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
    return 0;
}
extern "C" int elo_get_pdl(void) { return elo::pdl_enabled() ? 1 : 0; }
extern "C" long long elo_launch_count(void) { return elo::g_launches.load(std::memory_order_relaxed); }
