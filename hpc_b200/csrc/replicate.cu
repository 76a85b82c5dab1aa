// replicate.cu — sharded host I/O for the multi-GPU driver: B is replicated over NVLink, not N times over PCIe.
//
// No reference counterpart (the reference is single-GPU; SURVEY.md §8e: "each GPU gets ... full B replicated").
// One process per GPU. Every rank holds a full-size copy of B in symmetric (peer-mapped) memory. Per call, rank g
// uploads only ITS slice of B rows from the host, pushes that slice into every other rank's copy — one
// `multimem.st` per 16 bytes through the NVLS multicast address when there is one, else one `st.global` per peer —
// and a device-side flag barrier over the same peer mappings orders the pushes against the SpMM passes. PCIe then
// carries 4·b_rows·K/N bytes in and 4·rows_g·K bytes out per rank instead of the whole B on every rank.
#include "common.h"

namespace spmm_b200 {

namespace {

struct PtrTable {
    float *p[kMaxGather];
};
struct FlagTable {
    unsigned int *p[kMaxGather];
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid-stride float4 copy of this rank's slice into the peers' copies of B
__global__ void __launch_bounds__(256) push_rows_kernel(const float4 *__restrict__ src, long long off4, long long n4, int n,
                                                        PtrTable tg, int skip, float *mc) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcs(src + i);
        if (mc) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float4 *>(mc) + off4 + i),
                         "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
        } else {
            for (int t = 0; t < n; ++t)
                if (t != skip) reinterpret_cast<float4 *>(tg.p[t])[off4 + i] = v;
        }
    }
}

// one thread per rank: publish my arrival everywhere, then wait for everybody's in my own flag array.
// Epochs only grow (one per call and phase), compared wrap-safe.
__global__ void xrank_barrier_kernel(FlagTable fl, int world, int rank, int phase, unsigned int epoch) {
    const int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();   // whatever this stream wrote before (peer pushes of the previous kernel) is visible first
    st_release_sys(fl.p[t] + phase * world + rank, epoch);
    const unsigned int *mine = fl.p[rank] + phase * world + t;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) __nanosleep(64);
}

}  // namespace

int launch_push_rows(const float *src, long long off, long long count, int n, float *const *targets, int skip, float *mc,
                     cudaStream_t stream) {
    if (count <= 0) return 0;
    if ((count & 3) || (off & 3) || ((uintptr_t)src & 15)) {
        set_error("push_rows: slice must be 16-byte aligned and a multiple of 4 floats");
        return SPMM_B200_EINVAL;
    }
    PtrTable tg;
    for (int t = 0; t < kMaxGather; ++t) tg.p[t] = t < n ? targets[t] : nullptr;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        sms = 148;
    }
    const long long n4 = count / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 8ll * sms) blocks = 8ll * sms;
    push_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(src), off / 4, n4, n, tg, skip, mc);
    SB_CUDA(cudaGetLastError());
    return 0;
}

int launch_xrank_barrier(unsigned int *const *flags, int world, int rank, int phase, unsigned int epoch, cudaStream_t stream) {
    FlagTable fl;
    for (int t = 0; t < kMaxGather; ++t) fl.p[t] = t < world ? flags[t] : nullptr;
    xrank_barrier_kernel<<<1, 32, 0, stream>>>(fl, world, rank, phase, epoch);
    SB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace spmm_b200

using namespace spmm_b200;

extern "C" {

int spmm_b200_set_replicate(spmm_b200_t h, int world, int rank, float *const *peers_b, float *multicast_b,
                            unsigned int *const *peers_flags) {
    if (!h || world < 0 || world > kMaxGather || (world > 0 && (rank < 0 || rank >= world || !peers_b || !peers_flags))) {
        set_error("spmm_b200_set_replicate: bad arguments (world <= %d)", kMaxGather);
        return SPMM_B200_EINVAL;
    }
    for (int t = 0; t < world; ++t)
        if (!peers_b[t] || ((uintptr_t)peers_b[t] & 15) || !peers_flags[t]) {
            set_error("spmm_b200_set_replicate: buffer %d is null or not 16-byte aligned", t);
            return SPMM_B200_EINVAL;
        }
    h->rep_world = world;
    h->rep_rank = world > 0 ? rank : 0;
    for (int t = 0; t < kMaxGather; ++t) {
        h->rep_b[t] = t < world ? peers_b[t] : nullptr;
        h->rep_flags[t] = t < world ? peers_flags[t] : nullptr;
    }
    h->rep_mc = world > 0 ? multicast_b : nullptr;
    h->rep_epoch = 0;
    return 0;
}

int spmm_b200_run_host_sharded(spmm_b200_t h, const float *h_vin_rows, int row_begin, int row_count, float *h_vout,
                               void *stream) {
    if (!h || row_begin < 0 || row_count < 0 || (row_count > 0 && !h_vin_rows) || !h_vout) {
        set_error("spmm_b200_run_host_sharded: bad arguments");
        return SPMM_B200_EINVAL;
    }
    if (h->rep_world <= 0) {
        set_error("spmm_b200_run_host_sharded: spmm_b200_set_replicate has not been called");
        return SPMM_B200_ESTATE;
    }
    if (!h->plan.ready) {
        set_error("spmm_b200_run_host_sharded: preprocess has not been called");
        return SPMM_B200_ESTATE;
    }
    const int b_rows = h->b_rows > 0 ? h->b_rows : h->num_v;
    if ((long long)row_begin + row_count > b_rows || h->feat % 4 != 0) {
        set_error("spmm_b200_run_host_sharded: rows [%d, %d) outside B (%d rows), or feat_in %% 4 != 0", row_begin,
                  row_begin + row_count, b_rows);
        return SPMM_B200_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)h->num_v * h->feat;
    if (h->stage_elems < n) {
        cudaFree(h->d_stage_out);
        h->d_stage_out = nullptr;
        h->stage_elems = 0;
        if (n) SB_CUDA(cudaMalloc((void **)&h->d_stage_out, n * sizeof(float)));
        h->stage_elems = n;
    }
    float *local_b = h->rep_b[h->rep_rank];
    const long long off = (long long)row_begin * h->feat, cnt = (long long)row_count * h->feat;
    const unsigned int epoch = ++h->rep_epoch;
    int rc;
    // my slice: only my own passes read it, and those of the previous call are behind us on this stream
    if (cnt) SB_CUDA(cudaMemcpyAsync(local_b + off, h_vin_rows, (size_t)cnt * sizeof(float), cudaMemcpyHostToDevice, s));
    // phase 0: every rank is done reading its copy of B (previous call) before anybody overwrites a row of it
    if ((rc = launch_xrank_barrier(h->rep_flags, h->rep_world, h->rep_rank, 0, epoch, s))) return rc;
    if ((rc = launch_push_rows(local_b + off, off, cnt, h->rep_world, h->rep_b, h->rep_rank, h->rep_mc, s))) return rc;
    // phase 1: every rank's slice has landed everywhere
    if ((rc = launch_xrank_barrier(h->rep_flags, h->rep_world, h->rep_rank, 1, epoch, s))) return rc;
    if (n) {
        if ((rc = launch_spmm(h, local_b, h->d_stage_out, s, &h->plan.launches))) return rc;
        SB_CUDA(cudaMemcpyAsync(h_vout, h->d_stage_out, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    SB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"
