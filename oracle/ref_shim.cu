/*
 * oracle/ref_shim.cu — extern "C" doorway onto the UNMODIFIED reference (liblaf/hpc PA4).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as spmm_oracle.c). This file contains no SpMM code
 * of its own: it includes the reference's headers where they lie under
 * /root/reference/PA4/handout/include and is linked by oracle/Makefile with the
 * reference's own src/spmm_ref.cu, src/valid.cu, src/util.cu, src/data.cu and
 * src/spmm_cusparse.cu, compiled in place with the handout's flags
 * (-O3 --use_fast_math, PA4/handout/CMakeLists.txt:46) retargeted to sm_100a.
 * Output: oracle/_ref/libspmm_ref.so (git-ignored, travels to the GPU box).
 *
 * Callers: tests/ (GPU parity: reference kernel vs CPU restatement vs product) and
 * bench.py --impl reference (context timing of the reference's kernels).
 */
#include <cstring>
#include <string>

#include "spmm_ref.h"       // PA4/handout/include/spmm_ref.h:7-16
#include "spmm_cusparse.h"  // PA4/handout/include/spmm_cusparse.h:6-24
#include "valid.h"          // PA4/handout/include/valid.h:13-14
#include "data.h"           // PA4/handout/include/data.h:22

extern "C" {

/* SpMMRef(g, feat)->preprocess; ->run  (PA4/handout/test/test_spmm.cu:33-41). Synchronises. */
int ref_spmm_run(int *d_ptr, int *d_idx, float *d_val, float *d_vin, float *d_vout,
                 int num_v, int num_e, int feat) {
    CSR g(num_v, num_e, d_ptr, d_idx, d_val);
    SpMMRef op(&g, feat);
    op.preprocess(d_vin, d_vout);
    op.run(d_vin, d_vout);
    return (int)cudaDeviceSynchronize();
}

/* getAverageTimeWithWarmUp over SpMMRef::run (PA4/handout/include/util.h:141-151): seconds. */
double ref_spmm_time(int *d_ptr, int *d_idx, float *d_val, float *d_vin, float *d_vout,
                     int num_v, int num_e, int feat) {
    CSR g(num_v, num_e, d_ptr, d_idx, d_val);
    SpMMRef op(&g, feat);
    op.preprocess(d_vin, d_vout);
    return getAverageTimeWithWarmUp([&]() { op.run(d_vin, d_vout); });
}

/* SpMMCuSparse (PA4/handout/src/spmm_cusparse.cu:3-34) reads the globals kNumV/kNumE/kLen. */
int ref_cusparse_run(int *d_ptr, int *d_idx, float *d_val, float *d_vin, float *d_vout,
                     int num_v, int num_e, int feat, int timed, double *seconds) {
    kNumV = num_v; kNumE = num_e; kLen = feat;
    CSR g(num_v, num_e, d_ptr, d_idx, d_val);
    SpMMCuSparse op(&g, feat);
    op.preprocess(d_vin, d_vout);
    if (timed) {
        *seconds = getAverageTimeWithWarmUp([&]() { op.run(d_vin, d_vout); });
    } else {
        op.run(d_vin, d_vout);
    }
    return (int)cudaDeviceSynchronize();
}

/* valid(float*, float*, int) (PA4/handout/src/valid.cu:41-56): mismatch count. */
int ref_valid_float(float *d_y, float *d_y2, int num) { return valid(d_y, d_y2, num); }

/* load_graph (PA4/handout/src/data.cu:3-66); host only, usable without a GPU.
 * The reference `new[]`s the arrays; copy them out and release them here. */
int ref_load_graph(const char *datadir, const char *dset, int *num_v, int *num_e,
                   int *ptr_out, int *idx_out, long long ptr_cap, long long idx_cap) {
    basedir = std::string(datadir);
    if (basedir.empty() || basedir.back() != '/') basedir += "/";
    int nv = 0, ne = 0;
    int *p = NULL, *x = NULL;
    load_graph(std::string(dset), nv, ne, p, x);
    *num_v = nv; *num_e = ne;
    int rc = 0;
    if (ptr_out && idx_out) {
        if (ptr_cap < (long long)nv + 1 || idx_cap < ne) rc = -1;
        else {
            std::memcpy(ptr_out, p, sizeof(int) * ((size_t)nv + 1));
            std::memcpy(idx_out, x, sizeof(int) * (size_t)ne);
        }
    }
    delete[] p; delete[] x;
    return rc;
}

}  // extern "C"
