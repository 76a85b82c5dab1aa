// l2_persist_probe.cu — feasibility probe for "hot-row residency" on the products-shaped SpMM:
// would keeping the most-referenced B rows in a compact buffer under a persisting-L2 access-policy window cut the
// kernel's time? The probe replays the products shape's gather mix without the SpMM arithmetic:
//   50 % of the gathers from a window of +-8192 rows around the task's own row (the graph's local half),
//   50 % global with popularity ~ rank^-1/2 over 2.45 M rows — of which the top `hot` ranks are the candidates.
// Modes: A  every row gathered in place from B (2.5 GB)                          [what the engine does today]
//        B  the hot rows gathered from a compact buffer, no cache policy
//        C  compact buffer + persisting access-policy window on it (cudaLaunchAttributeAccessPolicyWindow)
// 1 KB rows, a warp per row, 4 rows in flight, tasks of 128 gathers in natural row order, grid-stride over tasks.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_persist_probe tools/l2_persist_probe.cu
#include <cuda_runtime.h>
#include <math.h>
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

struct Args {
    const float4 *b;      // M rows of 64 float4
    const float4 *hotbuf; // compact copy of the hot rows (mode B/C) or NULL (mode A)
    const int *rank_row;  // popularity rank -> row of B (a fixed scatter of the ranks over the id space)
    uint32_t m, hot, n_tasks;
    float *sink;
};

__global__ void __launch_bounds__(256) mix_kernel(Args a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t task = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; task < a.n_tasks; task += warps) {
        const uint32_t row0 = (uint32_t)(((uint64_t)task * a.m) / a.n_tasks);   // natural order: the task's own rows
        for (int it = 0; it < 32; ++it) {                                         // 32 x 4 = 128 gathers per task
            float4 v[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t h = hash32(task * 131u + it * 4u + u);
                const float4 *p;
                if (h & 1u) {   // local half
                    int r = (int)row0 + (int)(hash32(h) % 16384u) - 8192;
                    r = r < 0 ? 0 : (r >= (int)a.m ? (int)a.m - 1 : r);
                    p = a.b + (size_t)r * 64;
                } else {        // global half: rank = floor(m * u^2), u uniform  =>  P(rank <= k) = sqrt(k / m)
                    const float uu = (float)(hash32(h ^ 0x9e3779b9u) >> 8) * (1.0f / 16777216.0f);
                    const uint32_t rank = min(a.m - 1, (uint32_t)((float)a.m * uu * uu));
                    if (a.hotbuf && rank < a.hot) p = a.hotbuf + (size_t)rank * 64;
                    else p = a.b + (size_t)a.rank_row[rank] * 64;
                }
                v[u][0] = __ldg(p + lane);
                v[u][1] = __ldg(p + 32 + lane);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x += v[u][0].x + v[u][1].x;
                acc.y += v[u][0].y + v[u][1].y;
            }
        }
    }
    if (acc.x + acc.y == 123.456f) a.sink[0] = acc.x;
}

__global__ void fill_rank_row(int *rank_row, uint32_t m) {
    // a bijection rank -> row that scatters neighbouring ranks over the id space: multiply by an odd constant mod 2^k, cycle-walk
    uint32_t bits = 1;
    while ((1u << bits) < m) ++bits;
    const uint32_t mask = (1u << bits) - 1;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        uint32_t x = r;
        do {
            x = (x * 0x9E3779B1u + 0x7F4A7C15u) & mask;
        } while (x >= m);
        rank_row[r] = (int)x;
    }
}

__global__ void gather_hot(const float4 *b, const int *rank_row, uint32_t hot, float4 *hotbuf) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= hot) return;
    const float4 *src = b + (size_t)rank_row[w] * 64;
    hotbuf[(size_t)w * 64 + lane] = src[lane];
    hotbuf[(size_t)w * 64 + 32 + lane] = src[32 + lane];
}

static float run(const Args &a, int grid, bool window, size_t hot_bytes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cudaLaunchAttribute attr[1];
    if (window) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = (void *)a.hotbuf;
        attr[0].val.accessPolicyWindow.num_bytes = hot_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, mix_kernel, a));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    const uint32_t m = 2449029, n_tasks = 123718280u / 128u;
    int dev = 0, sms = 0, max_persist = 0, max_window = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    CK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    printf("{\"sms\": %d, \"max_persisting_l2\": %d, \"max_window\": %d}\n", sms, max_persist, max_window);
    float4 *b;
    int *rank_row;
    float *sink;
    CK(cudaMalloc((void **)&b, (size_t)m * 1024));
    CK(cudaMemset(b, 0, (size_t)m * 1024));
    CK(cudaMalloc((void **)&rank_row, sizeof(int) * m));
    CK(cudaMalloc((void **)&sink, 16));
    fill_rank_row<<<sms * 8, 256>>>(rank_row, m);
    CK(cudaDeviceSynchronize());
    const int grid = sms * 3;   // 24 warps per SM, as the K = 256 SpMM kernel
    Args a = {b, nullptr, rank_row, m, 0, n_tasks, sink};
    const float t_a = run(a, grid, false, 0);
    printf("{\"mode\": \"A in place\", \"ms\": %.4f}\n", t_a);
    for (uint32_t hot_mb : {16u, 32u, 48u, 64u, 72u}) {
        const uint32_t hot = hot_mb * 1024u;
        const size_t hot_bytes = (size_t)hot * 1024;
        float4 *hotbuf;
        CK(cudaMalloc((void **)&hotbuf, hot_bytes));
        gather_hot<<<(hot * 32 + 255) / 256, 256>>>(b, rank_row, hot, hotbuf);
        CK(cudaDeviceSynchronize());
        Args h = a;
        h.hotbuf = hotbuf;
        h.hot = hot;
        const float t_b = run(h, grid, false, 0);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
        const float t_c = run(h, grid, true, hot_bytes);
        CK(cudaCtxResetPersistingL2Cache());
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
        printf("{\"hot_mb\": %u, \"hot_share_of_global\": %.3f, \"ms_compact_no_policy\": %.4f, \"ms_compact_persisting\": %.4f, "
               "\"vs_in_place\": %.4f}\n",
               hot_mb, sqrt((double)hot / m), t_b, t_c, t_c / t_a);
        fflush(stdout);
        CK(cudaFree(hotbuf));
    }
    return 0;
}
