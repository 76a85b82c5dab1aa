// graph.cpp — host side of the graph path: synthetic generator, reference-format loader/writer,
// nnz-balanced row partition. No CUDA here.
//
// The reference loads real OGB/DGL graphs from outside its repository
// (PA4/handout/src/data.cu:3-66, PA4/handout/script/run_all.sh:11); they are not available
// offline, so the generator below produces graphs with the same (rows, nnz, max row nnz) —
// the only per-graph statistic the reference pins (PA4/workspace/phase_2.log). The definition
// is integer / correctly-rounded-double arithmetic only (mul, div, sqrt, floor), so the numpy
// restatement in oracle/graph_oracle.py reproduces it bit for bit.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "spmm_b200.h"

namespace spmm_b200 {
void set_error(const char *fmt, ...);
}
using spmm_b200::set_error;

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t stream_key(uint64_t seed, uint64_t stream) { return mix64(seed ^ mix64(stream)); }
inline uint64_t hash2(uint64_t key, uint64_t a, uint64_t b) { return mix64(mix64(key + a) ^ b); }
// (0, 1], exact in double
inline double unit_open0(uint64_t h) { return (double)((h >> 12) + 1) * (1.0 / 4503599627370496.0); }

constexpr uint64_t kStreamDegU = 1, kStreamDegZero = 2, kStreamCol = 3;

// weight u^(-k/4) from sqrt / mul / div only
inline double tail_weight(double u, int k) {
    const double s = sqrt(u);
    const double q = sqrt(s);
    switch (k) {
        case 1: return 1.0 / q;
        case 2: return 1.0 / s;
        case 3: return 1.0 / (s * q);
        default: return 1.0 / u;
    }
}

long long scaled_sum(const std::vector<double> &w, double s, int max_deg, int skip) {
    long long tot = 0;
    const long long n = (long long)w.size();
#pragma omp parallel for reduction(+ : tot)
    for (long long i = 0; i < n; ++i) {
        if (i == skip || w[i] == 0.0) continue;
        double d = floor(w[i] * s);
        if (d > (double)max_deg) d = (double)max_deg;
        tot += (long long)d;
    }
    return tot;
}

int gen_degrees_impl(int M, long long nnz, int max_deg, int tail_k, int zero_ppm, uint64_t seed, int *deg) {
    if (M <= 0 || nnz < 0 || max_deg < 0 || max_deg > M || tail_k < 1 || tail_k > 4 || zero_ppm < 0 ||
        zero_ppm > 1000000) {
        set_error("gen_degrees: bad arguments");
        return SPMM_B200_EINVAL;
    }
    if (nnz < max_deg) {
        set_error("gen_degrees: nnz < max_deg");
        return SPMM_B200_EINVAL;
    }
    const uint64_t ku = stream_key(seed, kStreamDegU), kz = stream_key(seed, kStreamDegZero);
    std::vector<double> w((size_t)M);
#pragma omp parallel for
    for (int i = 0; i < M; ++i) {
        const bool empty = (mix64(kz + (uint64_t)i) % 1000000ull) < (uint64_t)zero_ppm;
        w[i] = empty ? 0.0 : tail_weight(unit_open0(mix64(ku + (uint64_t)i)), tail_k);
    }
    // the heaviest row (lowest index among ties) is pinned to max_deg
    int top = 0;
    for (int i = 1; i < M; ++i)
        if (w[i] > w[top]) top = i;
    long long eligible = 0;
    for (int i = 0; i < M; ++i) eligible += (i != top && w[i] != 0.0);
    const long long want = nnz - max_deg;
    if (want > eligible * (long long)max_deg) {
        set_error("gen_degrees: nnz unreachable with max_deg");
        return SPMM_B200_EINVAL;
    }
    // largest scale with sum <= want, by a fixed number of bisection steps on [0, 2^40]
    double lo = 0.0, hi = 1099511627776.0;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        if (scaled_sum(w, mid, max_deg, top) <= want) lo = mid;
        else hi = mid;
    }
    long long tot = 0;
    for (int i = 0; i < M; ++i) {
        if (i == top) {
            deg[i] = max_deg;
        } else if (w[i] == 0.0) {
            deg[i] = 0;
        } else {
            double d = floor(w[i] * lo);
            if (d > (double)max_deg) d = (double)max_deg;
            deg[i] = (int)d;
            tot += deg[i];
        }
    }
    // hand the remainder out one nonzero at a time, in row order, wrapping as needed
    long long rem = want - tot;
    while (rem > 0) {
        long long given = 0;
        for (int i = 0; i < M && rem > 0; ++i) {
            if (i == top || w[i] == 0.0 || deg[i] >= max_deg) continue;
            ++deg[i];
            --rem;
            ++given;
        }
        if (given == 0) break;
    }
    if (rem != 0) {
        set_error("gen_degrees: could not place %lld nonzeros", rem);
        return SPMM_B200_EINVAL;
    }
    return 0;
}

// multiplier coprime to M that scatters popularity ranks over the id space
inline uint64_t scatter_mult(uint64_t M) {
    uint64_t a = (uint64_t)((double)M * 0.6180339887498949);
    if (a < 1) a = 1;
    auto gcd = [](uint64_t x, uint64_t y) {
        while (y) {
            uint64_t t = x % y;
            x = y;
            y = t;
        }
        return x;
    };
    while (gcd(a, M) != 1) ++a;
    return a;
}

// k-th candidate column of row r
inline int candidate(uint64_t key, uint64_t M, uint64_t mult, int local_ppm, int window, int r, uint64_t k) {
    const uint64_t h = hash2(key, (uint64_t)r, k);
    const uint64_t h2 = mix64(h);
    if ((h2 % 1000000ull) < (uint64_t)local_ppm) {
        const long long off = (long long)((h2 >> 20) % (uint64_t)(2 * (long long)window + 1)) - window;
        long long c = ((long long)r + off) % (long long)M;
        if (c < 0) c += (long long)M;
        return (int)c;
    }
    const double v = unit_open0(h);
    uint64_t rank = (uint64_t)floor((double)M * (v * v));   // popularity ~ rank^(-1/2)
    if (rank >= M) rank = M - 1;
    return (int)((rank * mult + 12345ull) % M);
}

}  // namespace

extern "C" int spmm_b200_gen_degrees(int num_v, long long nnz, int max_deg, int tail_k, int zero_ppm,
                                     uint64_t seed, int *deg) {
    if (!deg) {
        set_error("gen_degrees: null output");
        return SPMM_B200_EINVAL;
    }
    return gen_degrees_impl(num_v, nnz, max_deg, tail_k, zero_ppm, seed, deg);
}

extern "C" int spmm_b200_gen_graph(int num_v, long long nnz, int max_deg, int tail_k, int zero_ppm,
                                   int local_ppm, int window, uint64_t seed, int *ptr, int *idx) {
    if (!ptr || !idx || nnz > 0x7fffffffll || window < 0 || local_ppm < 0 || local_ppm > 1000000) {
        set_error("gen_graph: bad arguments");
        return SPMM_B200_EINVAL;
    }
    std::vector<int> deg((size_t)num_v > 0 ? (size_t)num_v : 1);
    int rc = gen_degrees_impl(num_v, nnz, max_deg, tail_k, zero_ppm, seed, deg.data());
    if (rc) return rc;
    const int M = num_v;
    ptr[0] = 0;
    for (int i = 0; i < M; ++i) ptr[i + 1] = ptr[i] + deg[i];
    const uint64_t key = stream_key(seed, kStreamCol);
    const uint64_t mult = scatter_mult((uint64_t)M);
#pragma omp parallel
    {
        std::vector<int> table;   // open addressing, -1 = empty
#pragma omp for schedule(dynamic, 64)
        for (int r = 0; r < M; ++r) {
            const int d = deg[r];
            if (d == 0) continue;
            int *out = idx + ptr[r];
            size_t cap = 16;
            while (cap < (size_t)d * 2) cap <<= 1;
            table.assign(cap, -1);
            int got = 0;
            const uint64_t max_draws = 64ull * (uint64_t)d + 64ull;
            for (uint64_t k = 0; got < d && k < max_draws; ++k) {
                const int c = candidate(key, (uint64_t)M, mult, local_ppm, window, r, k);
                size_t slot = (size_t)(mix64((uint64_t)c) & (cap - 1));
                bool dup = false;
                while (table[slot] != -1) {
                    if (table[slot] == c) {
                        dup = true;
                        break;
                    }
                    slot = (slot + 1) & (cap - 1);
                }
                if (dup) continue;
                table[slot] = c;
                out[got++] = c;
            }
            std::sort(out, out + got);
            if (got < d) {
                // dense row: top up with the smallest unused column ids
                std::vector<int> have(out, out + got);
                size_t hi = 0;
                for (int c = 0; got < d && c < M; ++c) {
                    while (hi < have.size() && have[hi] < c) ++hi;
                    if (hi < have.size() && have[hi] == c) continue;
                    out[got++] = c;
                }
                std::sort(out, out + got);
            }
        }
    }
    return 0;
}

// Host threads for the OpenMP parts of this library (graph generator). torchrun exports OMP_NUM_THREADS=1 to its
// workers; a caller that knows its share of the cores can override that here.
extern "C" int spmm_b200_set_host_threads(int n) {
    if (n < 1) {
        set_error("set_host_threads: need n >= 1");
        return SPMM_B200_EINVAL;
    }
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
    return 0;
}

// ---- reference graph files (PA4/handout/src/data.cu:3-66) -------------------------------------

static bool file_exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

extern "C" int spmm_b200_load_graph(const char *datadir, const char *dset, int *num_v, int *num_e, int *ptr,
                                    int *idx) {
    if (!datadir || !dset || !num_v || !num_e) {
        set_error("load_graph: null argument");
        return SPMM_B200_EINVAL;
    }
    std::string base(datadir);
    if (!base.empty() && base.back() != '/') base += "/";
    const std::string graph = base + dset + ".graph";
    const std::string ptrfile = graph + ".ptrdump", edgefile = graph + ".edgedump";
    const std::string config = base + dset + ".config";
    FILE *f = fopen(config.c_str(), "r");
    if (!f) {
        set_error("load_graph: cannot open %s", config.c_str());
        return SPMM_B200_EIO;
    }
    int nv = -1, ne = -1;
    const int got = fscanf(f, "%d %d", &nv, &ne);
    fclose(f);
    if (got != 2 || nv < 0 || ne < 0) {
        set_error("load_graph: malformed %s", config.c_str());
        return SPMM_B200_EIO;
    }
    *num_v = nv;
    *num_e = ne;
    if (!ptr && !idx) return 0;
    if (!ptr || !idx) {
        set_error("load_graph: ptr and idx must both be given");
        return SPMM_B200_EINVAL;
    }
    FILE *text = NULL;   // the text file holds ptr then idx, so it is read in order
    bool wrote_ptr = false;
    if (file_exists(ptrfile)) {
        FILE *fp = fopen(ptrfile.c_str(), "rb");
        const size_t n = fp ? fread(ptr, sizeof(int), (size_t)nv + 1, fp) : 0;
        if (fp) fclose(fp);
        if (n != (size_t)nv + 1) {
            set_error("load_graph: short read on %s", ptrfile.c_str());
            return SPMM_B200_EIO;
        }
    } else {
        text = fopen(graph.c_str(), "r");
        if (!text) {
            set_error("load_graph: cannot open %s", graph.c_str());
            return SPMM_B200_EIO;
        }
        for (int i = 0; i <= nv; ++i)
            if (fscanf(text, "%d", ptr + i) != 1) {
                fclose(text);
                set_error("load_graph: malformed ptr section in %s", graph.c_str());
                return SPMM_B200_EIO;
            }
        wrote_ptr = true;
    }
    if (ptr[nv] != ne) {   // data.cu:40-45
        if (text) fclose(text);
        set_error("load_graph: ptr[num_v] = %d but num_e = %d", ptr[nv], ne);
        return SPMM_B200_EIO;
    }
    if (wrote_ptr) {
        FILE *fp = fopen(ptrfile.c_str(), "wb");
        if (fp) {
            fwrite(ptr, sizeof(int), (size_t)nv + 1, fp);
            fclose(fp);
        }
    }
    if (file_exists(edgefile)) {
        if (text) fclose(text);
        FILE *fe = fopen(edgefile.c_str(), "rb");
        const size_t n = fe ? fread(idx, sizeof(int), (size_t)ne, fe) : 0;
        if (fe) fclose(fe);
        if (n != (size_t)ne) {
            set_error("load_graph: short read on %s", edgefile.c_str());
            return SPMM_B200_EIO;
        }
    } else {
        if (!text) {
            // ptr came from its dump: skip the ptr section of the text file
            text = fopen(graph.c_str(), "r");
            if (!text) {
                set_error("load_graph: cannot open %s", graph.c_str());
                return SPMM_B200_EIO;
            }
            int skip;
            for (int i = 0; i <= nv; ++i)
                if (fscanf(text, "%d", &skip) != 1) {
                    fclose(text);
                    set_error("load_graph: malformed ptr section in %s", graph.c_str());
                    return SPMM_B200_EIO;
                }
        }
        for (int i = 0; i < ne; ++i)
            if (fscanf(text, "%d", idx + i) != 1) {
                fclose(text);
                set_error("load_graph: malformed idx section in %s", graph.c_str());
                return SPMM_B200_EIO;
            }
        fclose(text);
        FILE *fe = fopen(edgefile.c_str(), "wb");
        if (fe) {
            fwrite(idx, sizeof(int), (size_t)ne, fe);
            fclose(fe);
        }
    }
    return 0;
}

extern "C" int spmm_b200_write_graph(const char *datadir, const char *dset, int num_v, int num_e,
                                     const int *ptr, const int *idx, int text) {
    if (!datadir || !dset || !ptr || (!idx && num_e > 0) || num_v < 0 || num_e < 0) {
        set_error("write_graph: bad arguments");
        return SPMM_B200_EINVAL;
    }
    std::string base(datadir);
    if (!base.empty() && base.back() != '/') base += "/";
    const std::string graph = base + dset + ".graph";
    FILE *f = fopen((base + dset + ".config").c_str(), "w");
    if (!f) {
        set_error("write_graph: cannot create config in %s", base.c_str());
        return SPMM_B200_EIO;
    }
    fprintf(f, "%d %d\n", num_v, num_e);
    fclose(f);
    if (text) {
        f = fopen(graph.c_str(), "w");
        if (!f) {
            set_error("write_graph: cannot create %s", graph.c_str());
            return SPMM_B200_EIO;
        }
        for (int i = 0; i <= num_v; ++i) fprintf(f, "%d ", ptr[i]);
        fprintf(f, "\n");
        for (int i = 0; i < num_e; ++i) fprintf(f, "%d ", idx[i]);
        fprintf(f, "\n");
        fclose(f);
    } else {
        f = fopen((graph + ".ptrdump").c_str(), "wb");
        if (!f) {
            set_error("write_graph: cannot create ptrdump");
            return SPMM_B200_EIO;
        }
        fwrite(ptr, sizeof(int), (size_t)num_v + 1, f);
        fclose(f);
        f = fopen((graph + ".edgedump").c_str(), "wb");
        if (!f) {
            set_error("write_graph: cannot create edgedump");
            return SPMM_B200_EIO;
        }
        fwrite(idx, sizeof(int), (size_t)num_e, f);
        fclose(f);
    }
    return 0;
}

// ---- multi-GPU partition (no reference counterpart) -----------------------------

extern "C" int spmm_b200_partition_rows(const int *h_ptr, int num_v, int parts, int *bounds) {
    if (!h_ptr || !bounds || num_v < 0 || parts <= 0) {
        set_error("partition_rows: bad arguments");
        return SPMM_B200_EINVAL;
    }
    const long long nnz = h_ptr[num_v];
    bounds[0] = 0;
    for (int g = 1; g < parts; ++g) {
        const long long target = (long long)g * nnz / parts;
        // first row r with ptr[r] >= target
        const int *p = std::lower_bound(h_ptr, h_ptr + num_v + 1, target,
                                        [](int a, long long t) { return (long long)a < t; });
        int r = (int)(p - h_ptr);
        if (r > num_v) r = num_v;
        if (r < bounds[g - 1]) r = bounds[g - 1];
        bounds[g] = r;
    }
    bounds[parts] = num_v;
    return 0;
}

// Cost-balanced variant: a row costs its nonzeros plus `row_cost` (the plan's per-row overhead: one header entry per
// column block, and the C row it reads and writes there). bounds[g] = first row r with ptr[r] + row_cost * r >=
// g * (nnz + row_cost * num_v) / parts, 64-bit integers; row_cost = 0 is spmm_b200_partition_rows.
extern "C" int spmm_b200_partition_rows_weighted(const int *h_ptr, int num_v, int parts, int row_cost, int *bounds) {
    if (!h_ptr || !bounds || num_v < 0 || parts <= 0 || row_cost < 0) {
        set_error("partition_rows_weighted: bad arguments");
        return SPMM_B200_EINVAL;
    }
    const long long total = (long long)h_ptr[num_v] + (long long)row_cost * num_v;
    bounds[0] = 0;
    for (int g = 1; g < parts; ++g) {
        const long long target = (long long)g * total / parts;
        int lo = 0, hi = num_v;   // first r in [0, num_v] with cost(r) >= target; cost is non-decreasing in r
        while (lo < hi) {
            const int mid = lo + (hi - lo) / 2;
            if ((long long)h_ptr[mid] + (long long)row_cost * mid < target) lo = mid + 1;
            else hi = mid;
        }
        bounds[g] = lo < bounds[g - 1] ? bounds[g - 1] : lo;
    }
    bounds[parts] = num_v;
    return 0;
}

extern "C" int spmm_b200_rebase_ptr(const int *h_ptr, int row_begin, int row_end, int *out_ptr) {
    if (!h_ptr || !out_ptr || row_begin < 0 || row_end < row_begin) {
        set_error("rebase_ptr: bad arguments");
        return SPMM_B200_EINVAL;
    }
    const int base = h_ptr[row_begin];
    for (int i = 0; i <= row_end - row_begin; ++i) out_ptr[i] = h_ptr[row_begin + i] - base;
    return 0;
}
