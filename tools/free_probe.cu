// free_probe.cu — what do cudaMalloc / cudaFree cost on this box, against a retaining cudaMemPool? (exploration for the plan
// allocator: destroying a reddit-sized plan took 0.4-0.7 s in cudaFree, profiles/r02_notes.md section 10)
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/free_probe tools/free_probe.cu ; ./tools/free_probe pool|malloc
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    cudaFree(0);
    const bool pool_mode = argc > 1 && !strcmp(argv[1], "pool");
    if (!pool_mode) {
        for (int rep = 0; rep < 2; ++rep)
            for (size_t mb : {64, 256, 1024, 2048}) {
                void *p;
                double t0 = now();
                cudaMalloc(&p, mb << 20);
                double t1 = now();
                cudaMemset(p, 1, mb << 20);
                cudaDeviceSynchronize();
                double t2 = now();
                cudaFree(p);
                double t3 = now();
                printf("{\"mode\": \"cudaMalloc\", \"rep\": %d, \"mb\": %zu, \"alloc_ms\": %.3f, \"free_ms\": %.3f}\n", rep, mb, t1 - t0, t3 - t2);
            }
        return 0;
    }
    cudaMemPool_t pool;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = 0;
    double t0 = now();
    cudaMemPoolCreate(&pool, &props);
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    printf("{\"mode\": \"pool\", \"create_ms\": %.3f}\n", now() - t0);
    for (int rep = 0; rep < 2; ++rep)
        for (size_t mb : {64, 256, 1024, 2048}) {
            void *p;
            double t0 = now();
            cudaError_t e = cudaMallocFromPoolAsync(&p, mb << 20, pool, 0);
            cudaStreamSynchronize(0);
            double t1 = now();
            cudaMemsetAsync(p, 1, mb << 20, 0);
            cudaStreamSynchronize(0);
            double t2 = now();
            cudaFreeAsync(p, 0);
            cudaStreamSynchronize(0);
            double t3 = now();
            printf("{\"mode\": \"pool\", \"rep\": %d, \"mb\": %zu, \"alloc_ms\": %.3f, \"free_ms\": %.3f, \"err\": \"%s\"}\n", rep, mb, t1 - t0, t3 - t2,
                   cudaGetErrorString(e));
        }
    t0 = now();
    cudaMemPoolTrimTo(pool, 0);
    printf("{\"mode\": \"pool\", \"trim_ms\": %.3f}\n", now() - t0);
    return 0;
}
