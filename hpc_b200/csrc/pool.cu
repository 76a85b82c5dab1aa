// pool.cu — device memory of the plans (and of the transposed operator's CSR) from a retaining CUDA memory pool.
//
// Takes the place of the cudaMalloc / cudaFree pairs of the student's SpMMOpt (PA4/workspace/src/spmm_opt.cu:57-66 allocate in
// preprocess, PA4/workspace/include/spmm_opt.h:18-20 frees in the destructor). Handing a reddit-sized plan (about 1 GB of
// panels) back to the driver with cudaFree took 0.4-0.7 s inside a process that also runs torch on some B200 boxes
// (profiles/r02_notes.md section 10) — per destroyed operator, and per preprocess call on a handle that already had a plan.
// The library therefore owns one cudaMemPool per device that keeps freed blocks (release threshold = max): the next
// plan takes them over in microseconds. spmm_b200_trim_memory() returns what nothing uses; SPMM_B200_POOL=0 in the
// environment selects plain cudaMalloc / cudaFree (decided once per device, so allocation and release always agree).
#include <stdlib.h>

#include <mutex>

#include "common.h"

namespace spmm_b200 {

namespace {
constexpr int kMaxPoolDevices = 64;
std::mutex g_pool_mu;
cudaMemPool_t g_pool[kMaxPoolDevices];
int g_pool_state[kMaxPoolDevices];   // 0 not tried yet, 1 ready, -1 unavailable (old driver, SPMM_B200_POOL=0)

cudaMemPool_t device_pool(int dev) {
    if (dev < 0 || dev >= kMaxPoolDevices) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_state[dev] == 0) {
        g_pool_state[dev] = -1;
        const char *env = getenv("SPMM_B200_POOL");
        int supported = 0;
        if (!(env && env[0] == '0') && cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess &&
            supported) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            unsigned long long keep = ~0ull;
            if (cudaMemPoolCreate(&g_pool[dev], &props) == cudaSuccess) {
                if (cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess)
                    g_pool_state[dev] = 1;
                else
                    cudaMemPoolDestroy(g_pool[dev]);
            }
        }
        cudaGetLastError();   // an unavailable pool is not an error of the call that asked
    }
    return g_pool_state[dev] == 1 ? g_pool[dev] : nullptr;
}

cudaMemPool_t current_pool() {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? device_pool(dev) : nullptr;
}
}   // namespace

// `bytes` on the current device, usable on `stream` from here on (and on any stream once `stream` was synchronised)
cudaError_t pool_alloc(void **ptr, size_t bytes, cudaStream_t stream) {
    if (cudaMemPool_t pool = current_pool()) return cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
    return cudaMalloc(ptr, bytes);
}

// a block of pool_alloc on the current device; the caller has made sure that whatever read it is ordered before `stream`
void pool_free(void *ptr, cudaStream_t stream) {
    if (!ptr) return;
    if (current_pool()) cudaFreeAsync(ptr, stream);
    else cudaFree(ptr);
}

int trim_pool_memory() {
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxPoolDevices) return 0;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_state[dev] != 1) return 0;
    SB_CUDA(cudaDeviceSynchronize());   // blocks freed in stream order become releasable once their stream got there
    SB_CUDA(cudaMemPoolTrimTo(g_pool[dev], 0));
    return 0;
}

}   // namespace spmm_b200
