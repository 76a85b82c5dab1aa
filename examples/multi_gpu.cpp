// examples/multi_gpu.cpp — the multi-GPU driver of the C ABI from a C++ host like the reference's harness
// (PA4/handout/test/main.cpp:5-24 is a single-GPU C++ program; the reference has no multi-GPU path, only the
// commented-out `// extern ncclComm_t* comms;` at PA4/handout/include/util.h:30).
//
// Build: g++ -std=c++17 -Iinclude -I/usr/local/cuda/include examples/multi_gpu.cpp -Lhpc_b200 -lspmm_b200
//            -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/hpc_b200 -o multi_gpu
// Run:   ./multi_gpu [n_devices]        (default: every visible device; one device listed twice works too)
//
// A 6-row graph times a 6x4 dense matrix: one call with host buffers (each device uploads its share of B and stores
// it into the other devices' copies over NVLink), then two stacked layers on the devices — with the NCCL all-gather
// of C and with the kernel's own epilogue — which must agree bit for bit.
#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "spmm_b200.h"

#define CK(x)                                                                   \
    do {                                                                        \
        if ((x) != 0) {                                                         \
            std::fprintf(stderr, "%s failed: %s\n", #x, spmm_b200_last_error()); \
            return 1;                                                           \
        }                                                                       \
    } while (0)

int main(int argc, char **argv) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) return 2;
    int n = argc > 1 ? std::atoi(argv[1]) : visible;
    if (n < 1) n = 1;
    std::vector<int> devices(n);
    for (int g = 0; g < n; ++g) devices[g] = g % visible;   // more "devices" than GPUs: the same GPU several times

    // A = ring + diagonal: row r = 1*B[r] + 2*B[(r+1) % 6]
    const int M = 6, K = 4;
    std::vector<int> ptr(M + 1), idx;
    std::vector<float> val;
    for (int r = 0; r < M; ++r) {
        ptr[r] = (int)idx.size();
        const int a = r, b = (r + 1) % M;
        idx.push_back(a < b ? a : b);
        idx.push_back(a < b ? b : a);
        val.push_back(a < b ? 1.f : 2.f);
        val.push_back(a < b ? 2.f : 1.f);
    }
    ptr[M] = (int)idx.size();
    std::vector<float> b(M * K), c(M * K, -1.f), want(M * K), want2(M * K);
    for (int i = 0; i < M * K; ++i) b[i] = (float)(i % 7);
    for (int r = 0; r < M; ++r)
        for (int j = 0; j < K; ++j) want[r * K + j] = b[r * K + j] + 2.f * b[((r + 1) % M) * K + j];
    for (int r = 0; r < M; ++r)
        for (int j = 0; j < K; ++j) want2[r * K + j] = want[r * K + j] + 2.f * want[((r + 1) % M) * K + j];

    spmm_b200_mg_t mg;
    CK(spmm_b200_mg_create(ptr.data(), idx.data(), val.data(), M, (int)idx.size(), K, n, devices.data(), &mg));
    CK(spmm_b200_mg_preprocess(mg));
    std::vector<int> bounds(n + 1);
    int peer = 0;
    CK(spmm_b200_mg_info(mg, nullptr, bounds.data(), &peer));
    std::printf("%d device(s), peer access %s, row blocks:", n, peer ? "on" : "off");
    for (int g = 0; g < n; ++g) std::printf(" [%d,%d)", bounds[g], bounds[g + 1]);
    std::printf("\n");

    CK(spmm_b200_mg_run_host(mg, b.data(), c.data()));   // B in, C out (host memory)
    if (std::memcmp(c.data(), want.data(), sizeof(float) * M * K) != 0) {
        std::fprintf(stderr, "run_host: wrong result\n");
        return 3;
    }
    for (int fused = 0; fused <= (peer || n == 1 ? 1 : 0); ++fused) {
        CK(spmm_b200_mg_set_fused(mg, fused));
        CK(spmm_b200_mg_run_host(mg, b.data(), c.data()));   // every device's B = the input again
        CK(spmm_b200_mg_run(mg));                            // layer 1
        if (!fused) CK(spmm_b200_mg_allgather(mg));
        CK(spmm_b200_mg_sync(mg));
        CK(spmm_b200_mg_swap(mg));                           // its C is layer 2's B
        CK(spmm_b200_mg_run(mg));                            // layer 2
        if (!fused) CK(spmm_b200_mg_allgather(mg));
        CK(spmm_b200_mg_sync(mg));
        for (int g = 0; g < n; ++g) {
            float *full = nullptr;
            CK(spmm_b200_mg_device_buffers(mg, g, nullptr, nullptr, &full, nullptr));
            cudaSetDevice(devices[g]);
            if (cudaMemcpy(c.data(), full, sizeof(float) * M * K, cudaMemcpyDeviceToHost) != cudaSuccess) return 4;
            if (std::memcmp(c.data(), want2.data(), sizeof(float) * M * K) != 0) {
                std::fprintf(stderr, "stacked layers (%s): wrong result on device %d\n", fused ? "fused" : "all-gather", g);
                return 5;
            }
        }
        CK(spmm_b200_mg_swap(mg));
        std::printf("two stacked layers, %s: ok\n", fused ? "kernel epilogue stores C rows to every device" : "NCCL all-gather-v");
    }
    CK(spmm_b200_mg_destroy(mg));
    return 0;
}
