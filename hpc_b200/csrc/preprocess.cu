// preprocess.cu — builds the execution plan of one (CSR, K) pair.
//
// Takes the place of the student's SpMMOpt::preprocess (PA4/workspace/src/spmm_opt.cu:37-69:
// copy ptr to the host, cut every row into <=256-nnz tasks, random_shuffle, upload, memset
// vout). The plan here is deterministic and fully specified so the CPU oracle
// (oracle/plan_oracle.py) can restate it and the tests compare bit for bit:
//
//   deg(r)    = ptr[r+1] - ptr[r]
//   heavy(r)  = deg(r) > seg_len
//   bucket(r) = bit length of deg(r)                      (0 for an empty row)
//   order     = rows by (bucket descending, r ascending)  (natural order when reorder = 0)
//   row_perm  = the non-heavy rows in that order
//   heavy rows, in that order, are cut into nseg = ceil(deg/seg_len) nnz-balanced segments
//               [begin + floor(j*deg/nseg), begin + floor((j+1)*deg/nseg)),  j = 0..nseg-1
//   panel     = the segments' {col * K/4, val} pairs back to back, each segment padded with nop entries {-1, 0} to a
//               multiple of 4 * (32 / lanes) entries (whole gather batches, 16-byte granules for the TMA copies)
//   lpanel    = the rows of row_perm as a stream: per row a header {0x80000000 | row, 0} and its {col, val}
//               entries, packed into equal-sized warp tasks by pack_light_host (rule stated there)
//   column blocks: when B exceeds the L2 (auto_col_blocks), every row is split at the band boundaries
//               (split[b][r]) and the whole plan above is built once per block over [split[b][r], split[b+1][r])
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "common.h"

namespace spmm_b200 {

// SPMM_B200_PREP_TRACE=1: per-phase wall times of preprocess on stderr (tools/prep_probe.py)
struct PrepTrace {
    bool on = getenv("SPMM_B200_PREP_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[prep] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

void free_plan(Plan &p) {
    for (BlockPlan &b : p.blocks) {
        cudaFree(b.d_row_perm);
        cudaFree(b.d_light_desc);
        cudaFree(b.d_ltask);
        cudaFree(b.d_utask);
        cudaFree(b.d_lpanel);
        cudaFree(b.d_heavy_rows);
        cudaFree(b.d_heavy_seg0);
        cudaFree(b.d_seg_desc);
        cudaFree(b.d_seg_hrow);
        cudaFree(b.d_seg_count);
        cudaFree(b.d_panel);
        cudaFree(b.d_part);
    }
    cudaFree(p.d_split);
    p = Plan();
}

static inline int bit_length(unsigned x) { return x ? 32 - __builtin_clz(x) : 0; }

int auto_seg_len(long long nnz, int lanes) {
    // Rows longer than this are cut into nnz-balanced segments, one warp each, their col/val
    // staged through shared memory by TMA. Measured on B200 (profiles/r01_sweep.md):
    //  * a full warp per row (K >= 128): equal-sized staged work units beat whole-row warps from
    //    ~256 nonzeros up on every graph shape;
    //  * narrower lane groups (K < 128): a segment carries only lanes*16 bytes per nonzero, so the
    //    per-segment partial + counter cost wants longer segments on big graphs (reddit K=32:
    //    1024), while small graphs are tail-bound and want short ones (arxiv K=32: 128):
    //    next power of two of nnz/65536, clamped to [128, 1024].
    if (lanes >= 32) return 256;
    long long t = nnz / 65536;
    int l = 128;
    while (l < t && l < 1024) l <<= 1;
    return l;
}

int auto_kslice(int num_v, int feat) {
    (void)num_v;
    if (feat % 4 != 0) return feat;
    // Measured on B200 (profiles/r01_sweep.md): narrower feature slices never paid for the extra
    // passes over col/val — the 126 MB L2 keeps the hot B rows even when B is larger than L2 —
    // so a pass covers up to 256 columns (32 lanes x 2 float4).
    return feat < 256 ? ((feat + 3) & ~3) : 256;
}

// Column blocks: when B does not fit the L2, A is processed as n column blocks (C = sum_b A[:, b]·B[b, :]),
// one pass each, so that every pass gathers from an L2-resident band of B and HBM sees B about once.
// Worth it only while a row still has enough nonzeros per block to amortise re-reading its C row.
int auto_col_blocks(long long b_rows, int feat, long long nnz, int num_v) {
    const long long b_bytes = b_rows * feat * 4ll;
    if (feat % 4 != 0 || num_v <= 0 || b_bytes <= (96ll << 20)) return 1;
    const long long band = 48ll << 20;
    long long nb = (b_bytes + band - 1) / band;
    if (nb > 16) return 1;
    if (nnz / num_v / nb < 64) return 1;
    return (int)nb;
}

// Host-only core of the plan (also exported as spmm_b200_plan_host for CPU-side tests).
// Row r owns CSR positions [rb[r], re[r]) (rb = ptr, re = ptr + 1 for the whole matrix; a column block
// passes its own bounds). skip_empty drops rows with no nonzero in the block (passes after the first).
// `pad`: every segment's panel span is padded (with nop entries) to a multiple of `pad` entries: 4 gather steps
// of all lane groups, so the kernel reads whole batches without bounds checks.
int plan_rows_host(const int *rb, const int *re, int M, int seg_len, int reorder, int skip_empty, int pad,
                   std::vector<int> &row_perm, std::vector<int> &heavy_rows, std::vector<int> &heavy_seg0,
                   std::vector<SegDesc> &segs, long long *panel_len_out) {
    // stable counting sort by bucket, descending
    std::vector<int> order((size_t)M);
    if (reorder) {
        size_t cnt[34] = {0};
        for (int r = 0; r < M; ++r) {
            const int d = re[r] - rb[r];
            if (d < 0) {
                set_error("CSR ptr decreases at row %d", r);
                return SPMM_B200_EINVAL;
            }
            ++cnt[bit_length((unsigned)d)];
        }
        size_t start[34];
        size_t run = 0;
        for (int b = 33; b >= 0; --b) {
            start[b] = run;
            run += cnt[b];
        }
        for (int r = 0; r < M; ++r) order[start[bit_length((unsigned)(re[r] - rb[r]))]++] = r;
    } else {
        for (int r = 0; r < M; ++r) order[r] = r;
    }

    row_perm.clear();
    heavy_rows.clear();
    heavy_seg0.clear();
    segs.clear();
    row_perm.reserve(M);
    long long panel_len = 0;
    for (int k = 0; k < M; ++k) {
        const int r = order[k];
        const int begin = rb[r];
        const int d = re[r] - begin;
        if (d == 0 && skip_empty) continue;
        if (d <= seg_len) {
            row_perm.push_back(r);
            continue;
        }
        heavy_rows.push_back(r);
        heavy_seg0.push_back((int)segs.size());
        const int nseg = (int)(((long long)d + seg_len - 1) / seg_len);
        for (int j = 0; j < nseg; ++j) {
            const int b = begin + (int)((long long)j * d / nseg);
            const int e = begin + (int)((long long)(j + 1) * d / nseg);
            SegDesc s;
            s.row = r;
            s.panel_off = (int)panel_len;
            s.len = e - b;
            s.nnz_begin = b;
            segs.push_back(s);
            panel_len += (s.len + pad - 1) / pad * pad;
        }
    }
    if (!heavy_rows.empty()) heavy_seg0.push_back((int)segs.size());
    if (panel_len > 0x7fffffffll) {
        set_error("panel too large");
        return SPMM_B200_EINVAL;
    }
    *panel_len_out = panel_len;
    return 0;
}

// Light rows as a stream. Rows are taken in plan order; each costs deg + 1 entries (a header carrying the row id,
// then its nonzeros). A task has `groups` lanes (one per lane group of the warp). A row goes to the least-filled
// lane of the current task (lowest index on ties); if that lane is non-empty and the row would take it past
// `steps` entries, the task is closed and the row opens the next one. A task's lanes are interleaved in the
// panel — entry j of lane g sits at off + j*groups + g — and padded with nops to the longest lane rounded up to a
// multiple of 4 steps (whole gather batches; it also keeps every task on a 16-byte boundary).
long long pack_light_host(const int *cost, int n, int groups, int steps, int *dst, std::vector<int2> &tasks) {
    tasks.clear();
    std::vector<int> fill((size_t)groups, 0);
    long long off = 0;
    auto close_task = [&]() {
        int mx = 0;
        for (int x : fill) mx = std::max(mx, x);
        mx = (mx + 3) & ~3;   // whole batches of 4 steps: the kernel reads a task without bounds checks
        tasks.push_back(make_int2((int)off, mx));
        off += (long long)mx * groups;
        std::fill(fill.begin(), fill.end(), 0);
    };
    for (int i = 0; i < n; ++i) {
        int g = 0;
        for (int q = 1; q < groups; ++q)
            if (fill[q] < fill[g]) g = q;
        if (fill[g] > 0 && fill[g] + cost[i] > steps) {
            close_task();
            g = 0;
        }
        dst[i] = (int)(off + (long long)fill[g] * groups + g);
        fill[g] += cost[i];
    }
    bool any = false;
    for (int x : fill) any |= x > 0;
    if (any) close_task();
    return off;
}

// Entries per lane group and stream task: 64 per warp task, at least 16 per lane group; 128 when a full warp
// serves each row and the block is longer than 64 waves (measured, profiles/r01_sweep.md: arxiv K=256 0.142 / 0.155 ms
// at 64 / 128; reddit K=256 6.28 / 6.14 ms; products K=256 12.45 / 12.09 ms; K=32 shapes flat from 8 to 64). Small graphs are quantised by waves — 1.1 waves of tasks
// take as long as 2 — so when the light rows amount to fewer than 4 waves of default-sized tasks the size is
// chosen to fill a whole number of waves of the warps the device keeps resident (`slots`, minus the heavy
// segments that run beside them).
int auto_light_steps(int groups, long long total_cost, long long slots, long long heavy_tasks) {
    int dflt = 64 / groups;
    if (dflt < 16) dflt = 16;
    if (slots <= 0 || total_cost <= 0) return dflt;
    if (groups == 1 && total_cost >= 64ll * slots * 64) return 128;   // very large graphs: fewer, longer tasks
    long long avail = slots - heavy_tasks % slots;
    if (avail < slots / 2) avail = slots;
    const long long per_wave = avail * groups * (long long)dflt;
    if (total_cost >= 4 * per_wave) return dflt;
    const long long waves = (total_cost + per_wave - 1) / per_wave;
    long long st = (total_cost + waves * avail * groups - 1) / (waves * avail * groups);
    st += st / 16 + 1;   // head-room: lanes are filled to within one row of the target
    if (st < 16) st = 16;
    return (int)(st < dflt ? st : dflt);
}

static int build_block(spmm_b200_handle *h, BlockPlan &bp, const int *rb, const int *re, int skip_empty,
                       cudaStream_t stream) {
    Plan &p = h->plan;
    const int M = h->num_v, K = h->feat;
    PrepTrace tr;
    std::vector<int> row_perm, heavy_rows, heavy_seg0;
    std::vector<SegDesc> segs;
    long long panel_len = 0;
    // Row order. Degree buckets (longest first) shorten the tail and keep the lanes of a task balanced; natural
    // order keeps neighbouring rows together, which lets a graph's locality hit in L2. Auto (-1): natural order
    // when a full warp serves each row (no lanes to balance) and the block is at least 8 waves of tasks long (no
    // tail to speak of) — measured 12.2 -> 11.3 ms on the products shape, neutral on reddit, and the opposite
    // (0.14 -> 0.19 ms) on the one-wave arxiv shape, which therefore keeps the buckets (profiles/r01_sweep.md).
    int reorder = (int)h->opt_reorder;
    if (reorder < 0) {
        long long total = 0;
        for (int r = 0; r < M; ++r) total += re[r] - rb[r] + 1;
        reorder = (p.lanes == 32 && p.slots > 0 && total >= 8ll * p.slots * 64) ? 0 : 1;
    }
    bp.reorder = reorder;
    const int pad = p.scalar ? 2 : 4 * (32 / p.lanes);
    int rc = plan_rows_host(rb, re, M, p.seg_len, reorder, skip_empty, pad, row_perm, heavy_rows, heavy_seg0,
                            segs, &panel_len);
    if (rc) return rc;
    tr.lap("block: plan_rows_host");
    bp.n_light = (int)row_perm.size();
    bp.n_heavy = (int)heavy_rows.size();
    bp.n_seg = (int)segs.size();
    bp.panel_len = panel_len;

    auto upload = [&](void **dst, const void *src, size_t bytes) -> int {
        if (bytes == 0) return 0;
        SB_CUDA(cudaMalloc(dst, bytes));
        SB_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return 0;
    };
    std::vector<int4> light((size_t)bp.n_light);
    std::vector<int> cost((size_t)bp.n_light), dst((size_t)bp.n_light);
    std::vector<int2> ltasks;
    const int groups = p.scalar ? 1 : 32 / p.lanes;
    for (int i = 0; i < bp.n_light; ++i) cost[i] = re[row_perm[i]] - rb[row_perm[i]] + 1;
    long long lpanel_len = 0;
    if (!p.scalar) {
        int steps = p.light_steps;
        if (h->opt_light_steps <= 0) {
            long long total = 0;
            for (int c : cost) total += c;
            steps = auto_light_steps(groups, total, p.slots, (long long)segs.size() * p.n_slices);
            if (&bp == &p.blocks[0]) p.light_steps = steps;   // reported by plan_info (block 0)
        }
        bp.light_steps = steps;
        lpanel_len = pack_light_host(cost.data(), bp.n_light, groups, steps, dst.data(), ltasks);
        if (lpanel_len > 0x7fffffffll) {
            set_error("light panel too large");
            return SPMM_B200_EINVAL;
        }
    }
    tr.lap("block: pack_light_host");
    for (int i = 0; i < bp.n_light; ++i) {
        const int r = row_perm[i];
        light[i] = make_int4(r, rb[r], re[r] - rb[r], p.scalar ? 0 : dst[i]);
    }
    bp.n_ltask = (int)ltasks.size();
    bp.lpanel_len = lpanel_len;

    // Scheduling order of the warp tasks. Bucketed rows: heavy segments (the longest rows) first, then the light
    // tasks. Natural order: light tasks and heavy segments merged by the row they start with, so that a heavy row
    // runs next to its neighbours (whose B rows it shares in L2) instead of ahead of everything.
    std::vector<int2> utask;
    utask.reserve(ltasks.size() + segs.size());
    if (reorder == 0 && !p.scalar) {
        size_t li = 0, first_i = 0;   // first_i: index in row_perm of the first row of light task li
        size_t si = 0;
        auto light_first_row = [&](size_t t) {
            while (first_i < (size_t)bp.n_light && dst[first_i] < ltasks[t].x) ++first_i;
            return first_i < (size_t)bp.n_light ? row_perm[first_i] : 0x7fffffff;
        };
        while (li < ltasks.size() || si < segs.size()) {
            const int lrow = li < ltasks.size() ? light_first_row(li) : 0x7fffffff;
            const int hrow = si < segs.size() ? segs[si].row : 0x7fffffff;
            if (si < segs.size() && hrow < lrow) {
                utask.push_back(make_int2(-1 - (int)si, 0));
                ++si;
            } else {
                utask.push_back(ltasks[li]);
                ++li;
            }
        }
    } else {
        for (size_t si = 0; si < segs.size(); ++si) utask.push_back(make_int2(-1 - (int)si, 0));
        for (const int2 &t : ltasks) utask.push_back(t);
    }
    bp.n_utask = (int)utask.size();
    tr.lap("block: light_desc + utask");
    std::vector<int> seg_hrow((size_t)bp.n_seg);
    for (int hr = 0; hr < bp.n_heavy; ++hr)
        for (int sgm = heavy_seg0[hr]; sgm < heavy_seg0[hr + 1]; ++sgm) seg_hrow[sgm] = hr;
    if ((rc = upload((void **)&bp.d_row_perm, row_perm.data(), sizeof(int) * row_perm.size()))) return rc;
    if ((rc = upload((void **)&bp.d_light_desc, light.data(), sizeof(int4) * light.size()))) return rc;
    if ((rc = upload((void **)&bp.d_utask, utask.data(), sizeof(int2) * utask.size()))) return rc;
    if (bp.n_ltask > 0) {
        if ((rc = upload((void **)&bp.d_ltask, ltasks.data(), sizeof(int2) * ltasks.size()))) return rc;
        SB_CUDA(cudaMalloc((void **)&bp.d_lpanel, sizeof(int2) * (size_t)lpanel_len));
        SB_CUDA(cudaMemsetAsync(bp.d_lpanel, 0xFF, sizeof(int2) * (size_t)lpanel_len, stream));   // nop entries
        if ((rc = launch_build_lpanel(bp.d_light_desc, bp.n_light, groups, K / 4, h->d_idx, h->d_val, bp.d_lpanel, stream)))
            return rc;
    }
    if (bp.n_heavy > 0) {
        if ((rc = upload((void **)&bp.d_seg_hrow, seg_hrow.data(), sizeof(int) * seg_hrow.size()))) return rc;
        const size_t ncnt = (size_t)bp.n_heavy * p.n_slices;
        SB_CUDA(cudaMalloc((void **)&bp.d_seg_count, sizeof(int) * ncnt));
        SB_CUDA(cudaMemsetAsync(bp.d_seg_count, 0, sizeof(int) * ncnt, stream));
        if ((rc = upload((void **)&bp.d_heavy_rows, heavy_rows.data(), sizeof(int) * heavy_rows.size()))) return rc;
        if ((rc = upload((void **)&bp.d_heavy_seg0, heavy_seg0.data(), sizeof(int) * heavy_seg0.size()))) return rc;
        if ((rc = upload((void **)&bp.d_seg_desc, segs.data(), sizeof(SegDesc) * segs.size()))) return rc;
        SB_CUDA(cudaMalloc((void **)&bp.d_panel, sizeof(int2) * (size_t)panel_len));
        SB_CUDA(cudaMalloc((void **)&bp.d_part, sizeof(float) * (size_t)bp.n_seg * K));
        if ((rc = launch_build_panel(bp.d_seg_desc, bp.n_seg, K / 4, pad, h->d_idx, h->d_val, bp.d_panel, stream))) return rc;
    }
    SB_CUDA(cudaStreamSynchronize(stream));   // host vectors go out of scope
    tr.lap("block: upload + panels");
    return 0;
}

int build_plan(spmm_b200_handle *h, cudaStream_t stream) {
    Plan &p = h->plan;
    free_plan(p);
    h->plan_select = 0;
    const int M = h->num_v, K = h->feat;
    const int b_rows = h->b_rows > 0 ? h->b_rows : M;
    if (K == 0 || M == 0) {
        // nothing to compute (launch_spmm returns before touching the plan): an empty, ready plan of one empty block
        p.block = (int)h->opt_block;
        p.n_col_blocks = 1;
        p.blocks.resize(1);
        p.ready = true;
        return 0;
    }
    p.block = (int)h->opt_block;
    p.scalar = (K % 4) != 0;
    p.kslice = h->opt_kslice > 0 ? (int)h->opt_kslice : auto_kslice(M, K);
    if (p.scalar) p.kslice = K;
    if (p.kslice > 256) p.kslice = 256;
    if (K > 0 && p.kslice > ((K + 3) & ~3)) p.kslice = (K + 3) & ~3;
    p.n_slices = (K > 0) ? (K + p.kslice - 1) / p.kslice : 0;
    if (!p.scalar && K > 0) shape_for_kslice(p.kslice, &p.lanes, &p.vec);
    p.seg_len = h->opt_seg_len > 0 ? (int)h->opt_seg_len : auto_seg_len(h->num_e, p.lanes);
    p.tune = (int)h->opt_tune;
    p.light_steps = h->opt_light_steps > 0 ? (int)h->opt_light_steps : 0;
    p.slots = p.scalar ? 0 : resident_warps(p.lanes, p.vec, p.tune, p.block);
    if (p.scalar) p.seg_len = 0x7fffffff;   // scalar fallback keeps every row whole
    if (!p.scalar && (long long)b_rows * (K / 4) > 0x7fffffffll) {
        set_error("B is too large for 32-bit row offsets: b_rows * feat_in / 4 = %lld >= 2^31", (long long)b_rows * (K / 4));
        return SPMM_B200_EINVAL;
    }
    int nb = h->opt_col_blocks > 0 ? (int)h->opt_col_blocks : auto_col_blocks(b_rows, K, h->num_e, M);
    if (p.scalar || M == 0 || nb < 1) nb = 1;
    if (nb > b_rows) nb = b_rows > 0 ? b_rows : 1;

    PrepTrace tr;
    std::vector<int> ptr((size_t)M + 1, 0);
    if (M > 0) {
        SB_CUDA(cudaMemcpyAsync(ptr.data(), h->d_ptr, sizeof(int) * ((size_t)M + 1), cudaMemcpyDeviceToHost,
                                stream));
        SB_CUDA(cudaStreamSynchronize(stream));
    }
    if (M > 0 && (ptr[0] != 0 || ptr[M] != h->num_e)) {
        set_error("CSR ptr is inconsistent: ptr[0]=%d ptr[num_v]=%d num_e=%d", ptr[0], ptr[M], h->num_e);
        return SPMM_B200_EINVAL;
    }
    for (int r = 0; r < M; ++r)   // whatever the row order: a negative degree would turn into out-of-bounds panel slots
        if (ptr[r + 1] < ptr[r]) {
            set_error("CSR ptr decreases at row %d", r);
            return SPMM_B200_EINVAL;
        }

    tr.lap("ptr D2H + checks");
    {   // columns must address rows of B
        int *d_bad = nullptr, bad = 0;
        SB_CUDA(cudaMalloc((void **)&d_bad, sizeof(int)));
        cudaError_t e = cudaMemsetAsync(d_bad, 0, sizeof(int), stream);
        int rc = e == cudaSuccess ? launch_check_cols(h->d_idx, h->num_e, b_rows, d_bad, stream) : 0;
        if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(stream);
        cudaFree(d_bad);
        if (rc) return rc;
        SB_CUDA(e);
        if (bad) {
            set_error("CSR idx holds a column outside [0, %d)", b_rows);
            return SPMM_B200_EINVAL;
        }
    }

    tr.lap("check_cols");
    // split every row at the column-block boundaries (needs ascending columns inside a row; a graph
    // that is not sorted falls back to a single block)
    std::vector<int> split;
    int cols_per_block = nb > 1 ? (b_rows + nb - 1) / nb : b_rows;
    if (nb > 1) nb = (b_rows + cols_per_block - 1) / cols_per_block;   // every band starts inside B (b_rows = 10, 7 bands -> 5 of 2 rows)
    if (nb > 1) {
        int *d_unsorted = nullptr;
        SB_CUDA(cudaMalloc((void **)&p.d_split, sizeof(int) * (size_t)(nb + 1) * M));
        SB_CUDA(cudaMalloc((void **)&d_unsorted, sizeof(int)));
        SB_CUDA(cudaMemsetAsync(d_unsorted, 0, sizeof(int), stream));
        int rc = launch_split_rows(h->d_ptr, h->d_idx, M, nb, cols_per_block, p.d_split, d_unsorted, stream);
        int unsorted = 0;
        cudaError_t e = cudaSuccess;
        if (rc == 0) e = cudaMemcpyAsync(&unsorted, d_unsorted, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (rc == 0 && e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(d_unsorted);
        if (rc) return rc;
        SB_CUDA(e);
        if (unsorted) {
            cudaFree(p.d_split);
            p.d_split = nullptr;
            nb = 1;
        } else {
            split.resize((size_t)(nb + 1) * M);
            SB_CUDA(cudaMemcpy(split.data(), p.d_split, sizeof(int) * split.size(), cudaMemcpyDeviceToHost));
        }
    }
    tr.lap("split_rows + D2H");
    p.n_col_blocks = nb;
    p.blocks.resize(nb);
    for (int b = 0; b < nb; ++b) {
        BlockPlan &bp = p.blocks[b];
        bp.col_begin = nb > 1 ? b * cols_per_block : 0;
        bp.col_end = nb > 1 ? (b + 1 == nb ? b_rows : (b + 1) * cols_per_block) : b_rows;
        const int *rb = nb > 1 ? split.data() + (size_t)b * M : ptr.data();
        const int *re = nb > 1 ? split.data() + (size_t)(b + 1) * M : ptr.data() + 1;
        // passes after the first skip rows without nonzeros in their band — except the last pass, which lists every
        // row: it is the one that delivers final rows (to C, to the stacked-layer targets, to run_host's host buffer)
        const bool skip_empty = b > 0 && b + 1 != nb;
        int rc = build_block(h, bp, rb, re, skip_empty, stream);
        if (rc) return rc;
    }
    p.ready = true;
    return 0;
}

// Re-stage the {col, val} panels from the caller's idx/val (asynchronous on `stream`): the plan's structure
// (row order, segments, tasks, slots) depends on ptr only, so new edge values need no new plan.
int refresh_panels(spmm_b200_handle *h, cudaStream_t stream) {
    Plan &p = h->plan;
    if (p.scalar || h->feat == 0 || h->num_v == 0) return 0;   // the scalar kernel reads idx/val at run time
    const int groups = 32 / p.lanes, pad = 4 * groups, k4 = h->feat / 4;
    for (BlockPlan &bp : p.blocks) {
        int rc;
        if (bp.n_ltask > 0 &&
            (rc = launch_build_lpanel(bp.d_light_desc, bp.n_light, groups, k4, h->d_idx, h->d_val, bp.d_lpanel, stream)))
            return rc;
        if (bp.n_seg > 0 && (rc = launch_build_panel(bp.d_seg_desc, bp.n_seg, k4, pad, h->d_idx, h->d_val, bp.d_panel, stream)))
            return rc;
    }
    return 0;
}

}  // namespace spmm_b200

// Host-only plan for CPU-side callers and tests (declared in include/spmm_b200.h).
extern "C" int spmm_b200_plan_host(const int *h_ptr, int num_v, int feat_in, long long seg_len, int reorder, int *row_perm,
                                   int *n_light, int *heavy_rows, int *n_heavy, int *heavy_seg0, int *seg_desc,
                                   int *n_seg, long long *panel_len) {
    using namespace spmm_b200;
    if (!h_ptr || num_v < 0 || !n_light || !n_heavy || !n_seg || !panel_len) {
        set_error("spmm_b200_plan_host: bad arguments");
        return SPMM_B200_EINVAL;
    }
    int lanes = 32, vec = 1;
    if (feat_in > 0 && feat_in % 4 == 0) shape_for_kslice(auto_kslice(num_v, feat_in), &lanes, &vec);
    if (seg_len <= 0) seg_len = (feat_in % 4 != 0) ? 0x7fffffffll : auto_seg_len(h_ptr[num_v], lanes);
    const int pad = (feat_in % 4 != 0) ? 2 : 4 * (32 / lanes);
    if (seg_len > 0x7fffffffll) seg_len = 0x7fffffffll;
    std::vector<int> rp, hr, hs;
    std::vector<SegDesc> segs;
    int rc = plan_rows_host(h_ptr, h_ptr + 1, num_v, (int)seg_len, reorder, 0, pad, rp, hr, hs, segs, panel_len);
    if (rc) return rc;
    *n_light = (int)rp.size();
    *n_heavy = (int)hr.size();
    *n_seg = (int)segs.size();
    if (row_perm) std::copy(rp.begin(), rp.end(), row_perm);
    if (heavy_rows) std::copy(hr.begin(), hr.end(), heavy_rows);
    if (heavy_seg0) std::copy(hs.begin(), hs.end(), heavy_seg0);
    if (seg_desc && !segs.empty()) memcpy(seg_desc, segs.data(), sizeof(SegDesc) * segs.size());
    return 0;
}

// Host-only light-stream packing (declared in include/spmm_b200.h).
extern "C" int spmm_b200_pack_light_host(const int *cost, int n, int groups, int steps, int *dst, int *ltask,
                                         int *n_ltask, long long *lpanel_len) {
    using namespace spmm_b200;
    if (n < 0 || groups < 1 || groups > 32 || (groups & (groups - 1)) || steps < 1 || !n_ltask || !lpanel_len ||
        (n > 0 && !cost)) {
        set_error("spmm_b200_pack_light_host: bad arguments");
        return SPMM_B200_EINVAL;
    }
    std::vector<int> d((size_t)n);
    std::vector<int2> tasks;
    *lpanel_len = pack_light_host(cost, n, groups, steps, d.data(), tasks);
    *n_ltask = (int)tasks.size();
    if (dst) std::copy(d.begin(), d.end(), dst);
    if (ltask && !tasks.empty()) memcpy(ltask, tasks.data(), sizeof(int2) * tasks.size());
    return 0;
}
