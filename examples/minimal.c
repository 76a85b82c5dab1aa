/* examples/minimal.c — the C boundary used from plain C (no C++, no torch): a 3x3 CSR times a 3x4 dense matrix.
 * Build:  gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/minimal.c -Lhpc_b200 -lspmm_b200 \
 *             -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/hpc_b200 -o minimal
 * Mirrors the reference's call sequence (PA4/handout/test/test_spmm.cu:33-41): create, preprocess once, run, sync. */
#include <cuda_runtime_api.h>
#include <stdio.h>

#include "spmm_b200.h"

#define CK(x)                                                         \
    do {                                                              \
        if ((x) != 0) {                                               \
            fprintf(stderr, "%s failed: %s\n", #x, spmm_b200_last_error()); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main(void) {
    /* row 0 = 2*B[1] + 3*B[2]; row 1 empty; row 2 = -1*B[0] */
    const int ptr[4] = {0, 2, 2, 3}, idx[3] = {1, 2, 0};
    const float val[3] = {2.f, 3.f, -1.f};
    float b[12], c[12];
    int *d_ptr, *d_idx;
    float *d_val, *d_b, *d_c;
    spmm_b200_t h;
    int i;
    for (i = 0; i < 12; ++i) b[i] = (float)i;
    if (cudaMalloc((void **)&d_ptr, sizeof ptr) || cudaMalloc((void **)&d_idx, sizeof idx) ||
        cudaMalloc((void **)&d_val, sizeof val) || cudaMalloc((void **)&d_b, sizeof b) || cudaMalloc((void **)&d_c, sizeof c))
        return 2;
    cudaMemcpy(d_ptr, ptr, sizeof ptr, cudaMemcpyHostToDevice);
    cudaMemcpy(d_idx, idx, sizeof idx, cudaMemcpyHostToDevice);
    cudaMemcpy(d_val, val, sizeof val, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b, sizeof b, cudaMemcpyHostToDevice);
    CK(spmm_b200_create(d_ptr, d_idx, d_val, 3, 3, 4, &h));
    CK(spmm_b200_preprocess(h, d_b, d_c, NULL));
    CK(spmm_b200_run(h, d_b, d_c, NULL));
    if (cudaDeviceSynchronize() != cudaSuccess) return 3;
    cudaMemcpy(c, d_c, sizeof c, cudaMemcpyDeviceToHost);
    for (i = 0; i < 3; ++i) printf("%g %g %g %g\n", c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]);
    CK(spmm_b200_destroy(h));
    /* expected: 32 37 42 47 / 0 0 0 0 / 0 -1 -2 -3 */
    return !(c[0] == 32.f && c[3] == 47.f && c[4] == 0.f && c[9] == -1.f && c[11] == -3.f);
}
