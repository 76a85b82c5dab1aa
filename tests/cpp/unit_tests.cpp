// unit_tests.cpp — gtest-free restatement of the reference's test driver, with SpMMB200 in the
// slot of SpMMOpt:
//   main / argParse           PA4/handout/test/main.cpp:5-24, src/util.cu:14-76  (--dataset --datadir --len)
//   SpMMTest fixture          PA4/handout/test/test_spmm.cu:8-29   (vin, vout, vout_ref, val: N(0,0.1))
//   validation                test_spmm.cu:31-44   (mismatches < M*K/10000 + 1, candidate first)
//   opt_performance           test_spmm.cu:55-62 + include/util.h:141-151 (10 warm-up + 20 timed)
// plus --shape <c0|arxiv|reddit|products> or --gen M,nnz,max_deg,tail_k,zero_ppm,local_ppm,window (with
// --dataset NAME for the log) to use the synthetic generator instead of files; tests/cpp/run_all.py sweeps the
// 13 dataset shapes of PA4/handout/script/run_all.sh:3-11 that way.
// TEST CODE: the reference output comes from the CPU oracle (oracle/liboracle.so), because the
// handout's SpMMRef is the reference's own CUDA kernel and is not part of this repository.
// Output keeps the dbg-macro form `[file:line (func)] time = <s> (double)` that
// PA4/workspace/plot.py:13-27 parses.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "spmm_b200.hpp"

extern "C" void oracle_spmm_f32(const int *ptr, const int *idx, const float *val, const float *vin, float *vout,
                                int feat, int row_begin, int row_end, int ftz, int nthreads);

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            std::fprintf(stderr, "Cuda failure: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(1);                                                               \
        }                                                                               \
    } while (0)

static double now() {
    return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

template <class F>
static double getAverageTimeWithWarmUp(F f) {   // util.h:141-151
    for (int i = 0; i < 10; ++i) f();
    double total = 0;
    for (int i = 0; i < 20; ++i) {
        CK(cudaDeviceSynchronize());
        double t0 = now();
        f();
        CK(cudaDeviceSynchronize());
        total += now() - t0;
    }
    return total / 20;
}

static float *allocate(long long num, unsigned long long stream_id) {   // data.h:24-37
    float *p = nullptr;
    long long n = (num + 511) / 512 * 512;
    CK(cudaMalloc((void **)&p, sizeof(float) * n));
    if (spmm_b200_fill_normal(p, n, 123, stream_id, 0.f, 0.1f, nullptr)) {
        std::fprintf(stderr, "%s\n", spmm_b200_last_error());
        std::exit(1);
    }
    return p;
}

int main(int argc, char **argv) {
    std::string dataset, datadir, shape, gen;
    int len = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--dataset")) dataset = argv[i + 1];
        else if (!std::strcmp(argv[i], "--datadir")) datadir = argv[i + 1];
        else if (!std::strcmp(argv[i], "--len")) len = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--shape")) shape = argv[i + 1];
        else if (!std::strcmp(argv[i], "--gen")) gen = argv[i + 1];
    }
    if (len <= 0 || (shape.empty() && gen.empty() && (dataset.empty() || datadir.empty()))) {
        std::fprintf(stderr, "usage: unit_tests (--dataset D --datadir DIR | --shape S) --len K\n");
        return 2;
    }
    int num_v = 0, num_e = 0;
    std::vector<int> ptr, idx;
    if (!gen.empty()) {
        long long m = 0, nnz = 0, mx = 0, k = 0, z = 0, l = 0, w = 0;
        if (std::sscanf(gen.c_str(), "%lld,%lld,%lld,%lld,%lld,%lld,%lld", &m, &nnz, &mx, &k, &z, &l, &w) != 7) {
            std::fprintf(stderr, "bad --gen spec\n");
            return 2;
        }
        num_v = (int)m; num_e = (int)nnz;
        ptr.resize(num_v + 1); idx.resize(num_e);
        if (spmm_b200_gen_graph((int)m, nnz, (int)mx, (int)k, (int)z, (int)l, (int)w, 123, ptr.data(), idx.data())) {
            std::fprintf(stderr, "%s\n", spmm_b200_last_error()); return 1;
        }
        if (dataset.empty()) dataset = "generated";
    } else if (!shape.empty()) {
        struct S { const char *n; int m; long long nnz; int mx, k, z, l, w; };
        const S shapes[] = {{"c0", 4096, 65536, 1024, 3, 50000, 300000, 64},
                            {"arxiv", 169343, 1166243, 13155, 3, 350000, 300000, 2048},
                            {"reddit", 232965, 114615892, 21657, 2, 0, 500000, 4096},
                            {"products", 2449029, 123718280, 17481, 2, 20000, 500000, 8192}};
        const S *s = nullptr;
        for (const S &c : shapes) if (shape == c.n) s = &c;
        if (!s) { std::fprintf(stderr, "unknown shape %s\n", shape.c_str()); return 2; }
        num_v = s->m; num_e = (int)s->nnz;
        ptr.resize(num_v + 1); idx.resize(num_e);
        if (spmm_b200_gen_graph(s->m, s->nnz, s->mx, s->k, s->z, s->l, s->w, 123, ptr.data(), idx.data())) {
            std::fprintf(stderr, "%s\n", spmm_b200_last_error()); return 1;
        }
        dataset = shape;
    } else {
        if (spmm_b200_load_graph(datadir.c_str(), dataset.c_str(), &num_v, &num_e, nullptr, nullptr)) {
            std::fprintf(stderr, "%s\n", spmm_b200_last_error()); return 1;
        }
        ptr.resize(num_v + 1); idx.resize(num_e);
        if (spmm_b200_load_graph(datadir.c_str(), dataset.c_str(), &num_v, &num_e, ptr.data(), idx.data())) {
            std::fprintf(stderr, "%s\n", spmm_b200_last_error()); return 1;
        }
    }
    std::fprintf(stderr, "[unit_tests.cpp:%d (main)] dset = \"%s\" (std::string)\n", __LINE__, dataset.c_str());
    std::fprintf(stderr, "[unit_tests.cpp:%d (main)] kLen = %d (int)\n", __LINE__, len);
    int *gptr, *gidx;                                             // main.cpp:11-16
    CK(cudaMalloc((void **)&gptr, sizeof(int) * (num_v + 1)));
    CK(cudaMalloc((void **)&gidx, sizeof(int) * (size_t)(num_e > 0 ? num_e : 1)));
    CK(cudaMemcpy(gptr, ptr.data(), sizeof(int) * (num_v + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gidx, idx.data(), sizeof(int) * (size_t)num_e, cudaMemcpyHostToDevice));

    const long long n = (long long)num_v * len;
    int failed = 0;
    std::printf("[==========] Running 2 tests from 1 test case.\n");
    // ---- SpMMTest.validation -------------------------------------------------------------------
    {
        std::printf("[ RUN      ] SpMMTest.validation\n");
        float *vin = allocate(n, 0), *vout = allocate(n, 1), *vout_ref = allocate(n, 2), *val = allocate(num_e, 3);
        CSR g(num_v, num_e, gptr, gidx, val);
        SpMMB200 *spmmer = new SpMMB200(&g, len);
        spmmer->preprocess(vin, vout);
        CK(cudaMemset(vout, 0, sizeof(float) * n));
        spmmer->run(vin, vout);
        CK(cudaDeviceSynchronize());
        std::vector<float> h_val(num_e), h_in(n), h_ref(n);
        CK(cudaMemcpy(h_val.data(), val, sizeof(float) * (size_t)num_e, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_in.data(), vin, sizeof(float) * n, cudaMemcpyDeviceToHost));
        oracle_spmm_f32(ptr.data(), idx.data(), h_val.data(), h_in.data(), h_ref.data(), len, 0, num_v, 0, 0);
        CK(cudaMemcpy(vout_ref, h_ref.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
        long long bad = -1;
        if (spmm_b200_valid(vout, vout_ref, n, &bad, nullptr)) { std::fprintf(stderr, "%s\n", spmm_b200_last_error()); return 1; }
        const bool ok = bad < n / 10000 + 1;                        // test_spmm.cu:43
        std::fprintf(stderr, "[unit_tests.cpp:%d (validation)] mismatches = %lld (long long)\n", __LINE__, bad);
        std::printf(ok ? "[       OK ] SpMMTest.validation\n" : "[  FAILED  ] SpMMTest.validation\n");
        failed += !ok;
        delete spmmer;
        cudaFree(vin); cudaFree(vout); cudaFree(vout_ref); cudaFree(val);
    }
    // ---- SpMMTest.opt_performance ----------------------------------------------------------------
    {
        std::printf("[ RUN      ] SpMMTest.opt_performance\n");
        float *vin = allocate(n, 4), *vout = allocate(n, 5), *val = allocate(num_e, 7);
        CSR g(num_v, num_e, gptr, gidx, val);
        SpMMB200 *spmmer = new SpMMB200(&g, len);
        spmmer->preprocess(vin, vout);
        double time = getAverageTimeWithWarmUp([&]() { spmmer->run(vin, vout); });
        std::fprintf(stderr, "[unit_tests.cpp:%d (TestBody)] time = %g (double)\n", __LINE__, time);
        std::printf("[       OK ] SpMMTest.opt_performance\n");
        delete spmmer;
        cudaFree(vin); cudaFree(vout); cudaFree(val);
    }
    std::printf(failed ? "[  FAILED  ] %d test.\n" : "[  PASSED  ] 2 tests.\n", failed);
    return failed ? 1 : 0;
}
