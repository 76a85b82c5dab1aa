// multi.cu — spmm_b200_mg_*: the multi-GPU driver behind the C ABI (one process, N devices).
//
// No reference counterpart: the reference is single-GPU (the only trace is the commented-out
// `// extern ncclComm_t* comms;` at PA4/handout/include/util.h:30). SURVEY.md §8e fixes the scheme: A's rows cut into
// contiguous blocks balanced by nnz, B replicated on every device, each device writes its own block of C; a collective
// only when a stacked layer needs the whole C everywhere.
//
//   run_host    device g uploads rows [g·M/N, (g+1)·M/N) of B over ITS PCIe link and stores them into every peer's copy of
//               B through NVLink peer mappings (push_rows_kernel), so PCIe carries B once in total, not once per device;
//               CUDA events order the pushes against the passes; every device downloads its own block of C.
//   set_fused   the SpMM kernel's row epilogue stores finished C rows into every device's full-size C (the next layer's B).
//   allgather   the baseline for the same step: NCCL all-gather-v (grouped ncclBroadcast, one root per row block).
//               libnccl is opened lazily with dlopen — the library has no link-time NCCL dependency.
#include <dlfcn.h>

#include <new>
#include <vector>

#include "common.h"

namespace spmm_b200 {

namespace {

// the handful of NCCL entry points the all-gather-v needs (nccl.h: ncclCommInitAll, ncclBroadcast, ...)
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(void **comms, int ndev, const int *devlist) = nullptr;
    int (*CommDestroy)(void *comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void *send, void *recv, size_t count, int dtype, int root, void *comm, cudaStream_t s) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    static constexpr int kFloat = 7;   // ncclFloat32
    bool open() {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast && GetErrorString;
    }
};

struct Dev {
    int id = 0;
    int row0 = 0, rows = 0;          // this device's block of A / C
    int up0 = 0, up_rows = 0;        // the rows of B this device uploads in run_host
    int *d_ptr = nullptr, *d_idx = nullptr;
    float *d_val = nullptr;
    float *d_b = nullptr;            // full B
    float *d_c = nullptr;            // local block of C
    float *d_full = nullptr;         // full C (fused epilogue / all-gather target), allocated on demand
    spmm_b200_t op = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t pushed = nullptr, done = nullptr;
    void *comm = nullptr;
};

}  // namespace

}  // namespace spmm_b200

using namespace spmm_b200;

struct spmm_b200_mg {
    int num_v = 0, num_e = 0, feat = 0, n = 0;
    std::vector<Dev> dev;
    std::vector<int> bounds;
    bool p2p = true, fused = false, ran = false, comms_ready = false, preprocessed = false;
    Nccl nccl;
};

namespace {

struct DeviceGuard {
    int saved = 0;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard() { cudaSetDevice(saved); }
};

int mg_fail(const char *what) {
    set_error("%s", what);
    return SPMM_B200_EINVAL;
}

int alloc_full(spmm_b200_mg *m) {
    for (Dev &d : m->dev) {
        if (d.d_full) continue;
        SB_CUDA(cudaSetDevice(d.id));
        SB_CUDA(cudaMalloc((void **)&d.d_full, sizeof(float) * std::max<size_t>(4, (size_t)m->num_v * m->feat)));
    }
    return 0;
}

int apply_gather(spmm_b200_mg *m) {
    std::vector<float *> targets;
    for (Dev &d : m->dev) targets.push_back(d.d_full);
    for (Dev &d : m->dev) {
        SB_CUDA(cudaSetDevice(d.id));
        int rc = m->fused ? spmm_b200_set_gather(d.op, m->n, targets.data(), nullptr, d.row0)
                          : spmm_b200_set_gather(d.op, 0, nullptr, nullptr, 0);
        if (rc) return rc;
        // with column blocks the mode is part of the plan (which rows the last pass lists): rebuild it
        if (m->preprocessed && !d.op->plan.ready && (rc = spmm_b200_preprocess(d.op, d.d_b, d.d_c, d.stream))) return rc;
    }
    return 0;
}

}  // namespace

extern "C" {

int spmm_b200_mg_create(const int *h_ptr, const int *h_idx, const float *h_val, int num_v, int num_e, int feat_in,
                        int n_devices, const int *devices, spmm_b200_mg_t *out) {
    if (!out || !h_ptr || num_v < 0 || num_e < 0 || feat_in < 0 || n_devices < 1 || n_devices > kMaxGather ||
        (num_e > 0 && (!h_idx || !h_val)))
        return mg_fail("spmm_b200_mg_create: bad arguments");
    // the partition below indexes idx / val by ptr: refuse an inconsistent ptr before anything is sliced
    if (h_ptr[0] != 0 || h_ptr[num_v] != num_e) {
        set_error("spmm_b200_mg_create: CSR ptr is inconsistent: ptr[0]=%d ptr[num_v]=%d num_e=%d", h_ptr[0], h_ptr[num_v], num_e);
        return SPMM_B200_EINVAL;
    }
    for (int r = 0; r < num_v; ++r)
        if (h_ptr[r + 1] < h_ptr[r]) {
            set_error("spmm_b200_mg_create: CSR ptr decreases at row %d", r);
            return SPMM_B200_EINVAL;
        }
    int visible = 0;
    SB_CUDA(cudaGetDeviceCount(&visible));
    for (int g = 0; g < n_devices; ++g) {
        const int id = devices ? devices[g] : g;
        if (id < 0 || id >= visible) {
            set_error("spmm_b200_mg_create: device %d is not visible (%d devices)", id, visible);
            return SPMM_B200_EINVAL;
        }
    }
    spmm_b200_mg *m = new (std::nothrow) spmm_b200_mg();
    if (!m) {
        set_error("spmm_b200_mg_create: out of host memory");
        return SPMM_B200_ENOMEM;
    }
    DeviceGuard guard;
    m->num_v = num_v;
    m->num_e = num_e;
    m->feat = feat_in;
    m->n = n_devices;
    m->dev.resize(n_devices);
    m->bounds.resize(n_devices + 1);
    int rc = spmm_b200_partition_rows(h_ptr, num_v, n_devices, m->bounds.data());
    auto bail = [&](int code) {
        spmm_b200_mg_destroy(m);
        return code;
    };
    if (rc) return bail(rc);
    for (int g = 0; g < n_devices; ++g) m->dev[g].id = devices ? devices[g] : g;
    // peer mappings: in one process a cudaMalloc pointer of device j is a valid address in kernels of device i once
    // peer access is on
    for (int i = 0; i < n_devices && n_devices > 1; ++i) {
        if (cudaSetDevice(m->dev[i].id) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__));
        for (int j = 0; j < n_devices; ++j) {
            if (i == j || m->dev[i].id == m->dev[j].id) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[i].id, m->dev[j].id);
            if (!can) {
                m->p2p = false;
                continue;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j].id, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->p2p = false;
            cudaGetLastError();
        }
    }
    std::vector<int> lptr;
    for (int g = 0; g < n_devices; ++g) {
        Dev &d = m->dev[g];
        d.row0 = m->bounds[g];
        d.rows = m->bounds[g + 1] - m->bounds[g];
        d.up0 = (int)((long long)num_v * g / n_devices);
        d.up_rows = (int)((long long)num_v * (g + 1) / n_devices) - d.up0;
        const int e0 = h_ptr[d.row0], e1 = h_ptr[d.row0 + d.rows];
        lptr.resize((size_t)d.rows + 1);
        if ((rc = spmm_b200_rebase_ptr(h_ptr, d.row0, d.row0 + d.rows, lptr.data()))) return bail(rc);
        cudaError_t e = cudaSetDevice(d.id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.pushed, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc((void **)&d.d_ptr, sizeof(int) * ((size_t)d.rows + 1));
        if (e == cudaSuccess) e = cudaMalloc((void **)&d.d_idx, sizeof(int) * std::max(1, e1 - e0));
        if (e == cudaSuccess) e = cudaMalloc((void **)&d.d_val, sizeof(float) * std::max(1, e1 - e0));
        if (e == cudaSuccess) e = cudaMalloc((void **)&d.d_b, sizeof(float) * std::max<size_t>(4, (size_t)num_v * feat_in));
        if (e == cudaSuccess) e = cudaMalloc((void **)&d.d_c, sizeof(float) * std::max<size_t>(4, (size_t)d.rows * feat_in));
        if (e == cudaSuccess) e = cudaMemcpy(d.d_ptr, lptr.data(), sizeof(int) * ((size_t)d.rows + 1), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && e1 > e0) e = cudaMemcpy(d.d_idx, h_idx + e0, sizeof(int) * (size_t)(e1 - e0), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && e1 > e0) e = cudaMemcpy(d.d_val, h_val + e0, sizeof(float) * (size_t)(e1 - e0), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return bail(cuda_fail(e, "spmm_b200_mg_create: device setup", __FILE__, __LINE__));
        if ((rc = spmm_b200_create(d.d_ptr, d.d_idx, d.d_val, d.rows, e1 - e0, feat_in, &d.op))) return bail(rc);
        if ((rc = spmm_b200_set_option(d.op, "b_rows", num_v))) return bail(rc);
    }
    *out = m;
    return 0;
}

int spmm_b200_mg_set_option(spmm_b200_mg_t m, const char *name, long long value) {
    if (!m) return mg_fail("spmm_b200_mg_set_option: null handle");
    for (Dev &d : m->dev) {
        int rc = spmm_b200_set_option(d.op, name, value);
        if (rc) return rc;
    }
    return 0;
}

int spmm_b200_mg_preprocess(spmm_b200_mg_t m) {
    if (!m) return mg_fail("spmm_b200_mg_preprocess: null handle");
    DeviceGuard guard;
    for (Dev &d : m->dev) {
        SB_CUDA(cudaSetDevice(d.id));
        int rc = spmm_b200_preprocess(d.op, d.d_b, d.d_c, d.stream);
        if (rc) return rc;
    }
    m->preprocessed = true;
    return 0;
}

int spmm_b200_mg_info(spmm_b200_mg_t m, int *n_devices, int *bounds, int *peer_access) {
    if (!m) return mg_fail("spmm_b200_mg_info: null handle");
    if (n_devices) *n_devices = m->n;
    if (bounds)
        for (int g = 0; g <= m->n; ++g) bounds[g] = m->bounds[g];
    if (peer_access) *peer_access = m->p2p ? 1 : 0;
    return 0;
}

int spmm_b200_mg_device_buffers(spmm_b200_mg_t m, int g, float **d_b, float **d_c_block, float **d_c_full, spmm_b200_t *op) {
    if (!m || g < 0 || g >= m->n) return mg_fail("spmm_b200_mg_device_buffers: bad arguments");
    if (d_b) *d_b = m->dev[g].d_b;
    if (d_c_block) *d_c_block = m->dev[g].d_c;
    if (d_c_full) *d_c_full = m->dev[g].d_full;
    if (op) *op = m->dev[g].op;
    return 0;
}

int spmm_b200_mg_set_fused(spmm_b200_mg_t m, int on) {
    if (!m) return mg_fail("spmm_b200_mg_set_fused: null handle");
    if (on && m->n > 1 && !m->p2p) {
        set_error("spmm_b200_mg_set_fused: the devices have no peer access to each other");
        return SPMM_B200_ESTATE;
    }
    if (on && m->feat % 4 != 0) return mg_fail("spmm_b200_mg_set_fused: needs feat_in % 4 == 0");
    DeviceGuard guard;
    int rc;
    if (on && (rc = alloc_full(m))) return rc;
    m->fused = on != 0;
    return apply_gather(m);
}

int spmm_b200_mg_run(spmm_b200_mg_t m) {
    if (!m) return mg_fail("spmm_b200_mg_run: null handle");
    DeviceGuard guard;
    // a fused run stores into the peers' full C: nobody may still be computing on it or copying out of it — the
    // caller's mg_sync between layers guarantees that; the launches below are independent of each other
    for (Dev &d : m->dev) {
        SB_CUDA(cudaSetDevice(d.id));
        int rc = spmm_b200_run(d.op, d.d_b, d.d_c, d.stream);
        if (rc) return rc;
        SB_CUDA(cudaEventRecord(d.done, d.stream));
    }
    m->ran = true;
    return 0;
}

int spmm_b200_mg_sync(spmm_b200_mg_t m) {
    if (!m) return mg_fail("spmm_b200_mg_sync: null handle");
    DeviceGuard guard;
    for (Dev &d : m->dev) {
        SB_CUDA(cudaSetDevice(d.id));
        SB_CUDA(cudaStreamSynchronize(d.stream));
    }
    return 0;
}

int spmm_b200_mg_swap(spmm_b200_mg_t m) {
    if (!m) return mg_fail("spmm_b200_mg_swap: null handle");
    DeviceGuard guard;
    int rc;
    if ((rc = alloc_full(m))) return rc;
    for (Dev &d : m->dev) std::swap(d.d_b, d.d_full);
    return apply_gather(m);
}

int spmm_b200_mg_run_host(spmm_b200_mg_t m, const float *h_vin, float *h_vout) {
    if (!m || ((size_t)m->num_v * m->feat > 0 && (!h_vin || !h_vout))) return mg_fail("spmm_b200_mg_run_host: bad arguments");
    DeviceGuard guard;
    const size_t K = (size_t)m->feat;
    const bool push = m->n > 1 && m->p2p && m->feat % 4 == 0;
    std::vector<float *> targets;
    for (Dev &d : m->dev) targets.push_back(d.d_b);
    // upload + replicate: device g brings in its slice of B rows and stores it into every copy of B
    for (int g = 0; g < m->n; ++g) {
        Dev &d = m->dev[g];
        SB_CUDA(cudaSetDevice(d.id));
        if (m->ran)   // nobody may still be gathering from a copy of B that is about to be overwritten
            for (Dev &o : m->dev) SB_CUDA(cudaStreamWaitEvent(d.stream, o.done, 0));
        const size_t off = (size_t)d.up0 * K, cnt = (size_t)d.up_rows * K;
        if (cnt) SB_CUDA(cudaMemcpyAsync(d.d_b + off, h_vin + off, cnt * sizeof(float), cudaMemcpyHostToDevice, d.stream));
        if (m->n > 1 && cnt) {
            if (push) {
                int rc = launch_push_rows(d.d_b + off, (long long)off, (long long)cnt, m->n, targets.data(), g, nullptr, d.stream);
                if (rc) return rc;
            } else {
                for (int o = 0; o < m->n; ++o)
                    if (o != g)
                        SB_CUDA(cudaMemcpyPeerAsync(m->dev[o].d_b + off, m->dev[o].id, d.d_b + off, d.id, cnt * sizeof(float), d.stream));
            }
        }
        SB_CUDA(cudaEventRecord(d.pushed, d.stream));
    }
    // passes + download of each device's block of C
    for (int g = 0; g < m->n; ++g) {
        Dev &d = m->dev[g];
        SB_CUDA(cudaSetDevice(d.id));
        for (int o = 0; o < m->n; ++o)
            if (o != g) SB_CUDA(cudaStreamWaitEvent(d.stream, m->dev[o].pushed, 0));
        int rc = spmm_b200_run(d.op, d.d_b, d.d_c, d.stream);
        if (rc) return rc;
        SB_CUDA(cudaEventRecord(d.done, d.stream));
        const size_t cnt = (size_t)d.rows * K;
        if (cnt) SB_CUDA(cudaMemcpyAsync(h_vout + (size_t)d.row0 * K, d.d_c, cnt * sizeof(float), cudaMemcpyDeviceToHost, d.stream));
    }
    m->ran = true;
    return spmm_b200_mg_sync(m);
}

int spmm_b200_mg_allgather(spmm_b200_mg_t m) {
    if (!m) return mg_fail("spmm_b200_mg_allgather: null handle");
    DeviceGuard guard;
    int rc;
    if ((rc = alloc_full(m))) return rc;
    const size_t K = (size_t)m->feat;
    bool distinct = true;
    for (int i = 0; i < m->n; ++i)
        for (int j = 0; j < i; ++j) distinct &= m->dev[i].id != m->dev[j].id;
    if (m->n == 1 || !distinct) {
        // one device (possibly listed several times): NCCL has nothing to do / refuses duplicate devices — plain copies,
        // each ordered after its source block's run
        for (Dev &d : m->dev) {
            SB_CUDA(cudaSetDevice(d.id));
            for (Dev &o : m->dev) {
                if (m->ran) SB_CUDA(cudaStreamWaitEvent(d.stream, o.done, 0));
                if ((size_t)o.rows * K > 0)
                    SB_CUDA(cudaMemcpyAsync(d.d_full + (size_t)o.row0 * K, o.d_c, sizeof(float) * o.rows * K, cudaMemcpyDeviceToDevice,
                                            d.stream));
            }
        }
        return 0;
    }
    if (!m->nccl.open()) {
        set_error("spmm_b200_mg_allgather: libnccl.so.2 could not be opened (%s)", dlerror());
        return SPMM_B200_ESTATE;
    }
    if (!m->comms_ready) {
        std::vector<void *> comms(m->n);
        std::vector<int> ids;
        for (Dev &d : m->dev) ids.push_back(d.id);
        int e = m->nccl.CommInitAll(comms.data(), m->n, ids.data());
        if (e) {
            set_error("ncclCommInitAll: %s", m->nccl.GetErrorString(e));
            return SPMM_B200_ESTATE;
        }
        for (int g = 0; g < m->n; ++g) m->dev[g].comm = comms[g];
        m->comms_ready = true;
    }
    // all-gather-v: row blocks differ in size (the partition balances nnz, not rows), so one broadcast per block,
    // grouped into a single NCCL launch per device
    int e = m->nccl.GroupStart();
    for (int g = 0; g < m->n && !e; ++g) {
        Dev &d = m->dev[g];
        for (int r = 0; r < m->n && !e; ++r) {
            const size_t cnt = (size_t)m->dev[r].rows * K;
            if (!cnt) continue;
            e = m->nccl.Broadcast(d.d_c, d.d_full + (size_t)m->dev[r].row0 * K, cnt, Nccl::kFloat, r, d.comm, d.stream);
        }
    }
    int e2 = m->nccl.GroupEnd();
    if (e || e2) {
        set_error("spmm_b200_mg_allgather: %s", m->nccl.GetErrorString(e ? e : e2));
        return SPMM_B200_ESTATE;
    }
    return 0;
}

int spmm_b200_mg_destroy(spmm_b200_mg_t m) {
    if (!m) return 0;
    DeviceGuard guard;
    for (Dev &d : m->dev) {
        cudaSetDevice(d.id);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.comm && m->nccl.CommDestroy) m->nccl.CommDestroy(d.comm);
        if (d.op) spmm_b200_destroy(d.op);
        cudaFree(d.d_ptr);
        cudaFree(d.d_idx);
        cudaFree(d.d_val);
        cudaFree(d.d_b);
        cudaFree(d.d_c);
        cudaFree(d.d_full);
        if (d.pushed) cudaEventDestroy(d.pushed);
        if (d.done) cudaEventDestroy(d.done);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    cudaGetLastError();
    delete m;
    return 0;
}

}  // extern "C"
