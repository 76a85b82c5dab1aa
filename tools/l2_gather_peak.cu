// l2_gather_peak.cu — measures the ceiling the K=256 SpMM gather path divides by: random whole-row gathers
// (row = 1 KB / 512 B / 128 B, 128-bit loads, a lane group per row exactly like spmm_kernel) from a buffer that is
// L2-resident (48 MB, one column-block band) or far larger than the L2 (4 GB), sweeping resident warps per SM and
// gathers in flight per lane group. No arithmetic beyond one add per float4 (keeps the loads alive), no index
// traffic (row ids come from a counter hash), so the rate is what the L2 -> SM fabric (or HBM) delivers to this
// access pattern. SURVEY.md §8d: "to be measured by the builder with a gather micro-benchmark".
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_gather_peak tools/l2_gather_peak.cu
//   tools/l2_gather_peak > profiles/r02_l2_gather_peak.jsonl
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                            \
    do {                                                                                 \
        cudaError_t e = (x);                                                             \
        if (e != cudaSuccess) {                                                          \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                     \
        }                                                                                \
    } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

// LANES lanes per row, VEC float4 per lane (row bytes = LANES * VEC * 16), U rows in flight per lane group
template <int LANES, int VEC, int U>
__global__ void __launch_bounds__(256) gather_kernel(const float4 *__restrict__ buf, uint32_t n_rows, int iters, float *sink) {
    constexpr int GROUPS = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int l = lane % LANES, g = lane / LANES;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t ctr = gw * 0x9E3779B9u + g * 0x85EBCA6Bu;
    for (int it = 0; it < iters; ++it) {
        float4 b[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t row = (uint32_t)(((uint64_t)hash32(ctr + u) * n_rows) >> 32);
            const float4 *p = buf + (size_t)row * (LANES * VEC) + l;
#pragma unroll
            for (int v = 0; v < VEC; ++v) b[u][v] = __ldg(p + v * LANES);
        }
        ctr += U;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc.x += b[u][v].x;
                acc.y += b[u][v].y;
                acc.z += b[u][v].z;
                acc.w += b[u][v].w;
            }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;   // never true: keeps the loads
    (void)GROUPS;
}

template <int LANES, int VEC, int U>
void run_case(const float4 *buf, size_t bytes, int sms, int ctas_per_sm, float *sink, const char *where) {
    const uint32_t row_f4 = LANES * VEC;
    const uint32_t n_rows = (uint32_t)(bytes / (row_f4 * 16));
    int maxb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, gather_kernel<LANES, VEC, U>, 256, 0));
    if (ctas_per_sm > maxb) return;
    const int grid = sms * ctas_per_sm;
    // ~32 GB of gathers per timed launch at full residency
    const double target = 32e9;
    const double per_iter = (double)grid * 8 * (32 / LANES) * U * row_f4 * 16;
    int iters = (int)(target / per_iter);
    if (iters < 16) iters = 16;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    gather_kernel<LANES, VEC, U><<<grid, 256>>>(buf, n_rows, iters / 4 + 1, sink);   // warm-up, fills the L2
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        gather_kernel<LANES, VEC, U><<<grid, 256>>>(buf, n_rows, iters, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    const double gbs = per_iter * iters / (best * 1e-3) / 1e9;
    printf("{\"buffer\": \"%s\", \"buffer_mb\": %.0f, \"row_bytes\": %u, \"lanes\": %d, \"vec\": %d, \"in_flight\": %d, "
           "\"warps_per_sm\": %d, \"ms\": %.4f, \"gather_gbs\": %.1f}\n",
           where, bytes / 1048576.0, row_f4 * 16, LANES, VEC, U, ctas_per_sm * 8, best, gbs);
    fflush(stdout);
    CK(cudaEventDestroy(a));
    CK(cudaEventDestroy(b));
}

template <int LANES, int VEC>
void sweep(const float4 *buf, size_t bytes, int sms, float *sink, const char *where) {
    for (int ctas : {1, 2, 3, 4, 6, 8}) {
        run_case<LANES, VEC, 2>(buf, bytes, sms, ctas, sink, where);
        run_case<LANES, VEC, 4>(buf, bytes, sms, ctas, sink, where);
        run_case<LANES, VEC, 8>(buf, bytes, sms, ctas, sink, where);
    }
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_khz\": %d, \"l2_bytes\": %d}\n", prop.name, sms, clk, prop.l2CacheSize);
    const size_t small = 48ull << 20, big = 4ull << 30;
    float4 *buf = nullptr;
    float *sink = nullptr;
    CK(cudaMalloc((void **)&buf, big));
    CK(cudaMalloc((void **)&sink, 16));
    CK(cudaMemset(buf, 0, big));
    // K = 256 (1 KB rows: 32 lanes x 2 float4), K = 128 (512 B: 32 x 1), K = 32 (128 B: 8 lanes x 1, 4 rows per warp load)
    sweep<32, 2>(buf, small, sms, sink, "l2_resident");
    sweep<32, 1>(buf, small, sms, sink, "l2_resident");
    sweep<8, 1>(buf, small, sms, sink, "l2_resident");
    sweep<32, 2>(buf, big, sms, sink, "hbm");
    CK(cudaFree(buf));
    CK(cudaFree(sink));
    return 0;
}
