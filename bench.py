#!/usr/bin/env python
"""bench.py — CSR SpMM throughput on B200 (BASELINE.json: "SpMM GFLOP/s & HBM GB/s (% roofline)
at K=32/256, 1/2/4/8 B200 vs host CPU").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload reddit_k256] [--impl reference]

A step is one SpMM (C = A·B) of the named synthetic graph: every kernel `spmm_b200_run`
launches, through the C ABI. For N > 1 (torchrun) A's rows are partitioned by nnz over the ranks
(B replicated, no data-path collective), the timed region is bracketed by a barrier and a device
synchronise, the per-rank device time is taken with CUDA events and the MAX over ranks is
reported; `value` is the whole job's GFLOP/s = 2·nnz·K / step time ("strong" scaling: the graph
is fixed, each rank gets 1/N of its nonzeros).

`--impl reference` times the CPU restatement of the reference's SpMM (oracle/spmm_oracle.c,
OpenMP over rows — the reference itself has no CPU path: PA4/handout/src/spmm_ref.cu is a CUDA
kernel) on the box's host cores, on a bounded row sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (graph shape, K)
    "c0_k32": ("c0", 32),
    "arxiv_k32": ("arxiv", 32),
    "arxiv_k256": ("arxiv", 256),
    "reddit_k32": ("reddit", 32),
    "reddit_k256": ("reddit", 256),
    "products_k256": ("products", 256),
}
SEED = 123
L2_BYTES = 126 << 20


def bytes_min(m, nnz, k, b_rows=None):
    """ptr + idx + val + B read once + C written once."""
    b_rows = m if b_rows is None else b_rows
    return 4 * (m + 1) + 8 * nnz + 4 * b_rows * k + 4 * m * k


def bytes_gather(m, nnz, k):
    return 4 * (m + 1) + 8 * nnz + 4 * nnz * k + 4 * m * k


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (pynvml, 50 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_sample(ptr, idx, k, seconds, nthreads=0):
    """Time the CPU oracle on a bounded prefix of rows; returns (gflops, description, cores)."""
    from oracle import cpu as O
    m, nnz = len(ptr) - 1, int(ptr[-1])
    val = O.fill_normal(nnz, SEED, 1)
    b = O.fill_normal(m * k, SEED, 2)
    out = np.zeros(m * k, np.float32)
    # all the cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would hide them)
    cores = nthreads or max(O.num_threads(), len(os.sched_getaffinity(0)))

    def run(rows):
        t = time.perf_counter()
        O.spmm_f32(ptr, idx, val, b, k, 0, rows, False, cores, out)
        return time.perf_counter() - t

    # calibrate on ~1% of the nonzeros, then size the sample for `seconds`: a row prefix when the
    # whole graph would take longer, otherwise the whole graph repeated
    r1 = int(np.searchsorted(ptr, max(1, nnz // 100), side="left"))
    r1 = max(1, min(m, r1))
    run(r1)
    t1 = run(r1)
    rate = 2.0 * int(ptr[r1]) * k / max(t1, 1e-9)
    want_nnz = min(nnz, int(rate * seconds / (2.0 * k)))
    rows = int(np.searchsorted(ptr, want_nnz, side="left"))
    rows = max(r1, min(m, rows))
    snnz = int(ptr[rows])
    reps, t = 0, 0.0
    while reps == 0 or (rows == m and t < seconds and reps < 10000):
        t += run(rows)
        reps += 1
    desc = f"rows [0,{rows}) of {m} ({snnz} of {nnz} nnz) x {reps} passes, {t:.2f} s"
    return 2.0 * snnz * k * reps / t / 1e9, desc, cores, t / reps, snnz


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import hpc_b200 as H  # host-side graph generator only (no GPU work on this arm)
    H.set_host_threads(len(os.sched_getaffinity(0)))
    shape, k = WORKLOADS[args.workload]
    ptr, idx = H.gen_named_graph(shape, SEED)
    m, nnz = len(ptr) - 1, len(idx)
    per_step = max(1.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, desc, cores = [], "", 0
    for s in range(args.warmup + args.steps):
        g, desc, cores, t, snnz = cpu_sample(ptr, idx, k, per_step)
        if s >= args.warmup:
            vals.append((g, t, snnz))
    gf = float(np.mean([v[0] for v in vals]))
    ms_full = 2.0 * nnz * k / (gf * 1e9) * 1e3
    line = {
        "impl": "reference", "metric": "spmm_gflops", "value": round(gf, 3), "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_full, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "graph": shape, "num_v": m, "nnz": nnz, "K": k,
                   "note": "ms_per_step extrapolates the sampled rate to the full graph"},
        "cpu_baseline": {"value": round(gf, 3), "unit": "GFLOP/s", "cores": cores, "kind": "port",
                         "sample": desc + " per step (OpenMP restatement of spmm_ref.cu:3-17; the reference has no CPU path)"},
        "e2e": {"value": round(gf, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def quick_measure(H, torch, shape, k, dev, steps=10, warmup=3):
    """Kernel-only numbers for one more BASELINE config on the same GPU (L2 flushed between iterations)."""
    ptr, idx = H.gen_named_graph(shape, SEED)
    m, nnz = len(ptr) - 1, len(idx)
    g = H.CSR(m, nnz, torch.from_numpy(ptr).to(dev), torch.from_numpy(idx).to(dev),
              H.fill_normal(torch.empty(nnz, dtype=torch.float32, device=dev), SEED, 1))
    vin = H.fill_normal(torch.empty(m * k, dtype=torch.float32, device=dev), SEED, 2)
    vout = torch.empty(m * k, dtype=torch.float32, device=dev)
    op = H.SpMMB200(g, k)
    op.preprocess(vin, vout)
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        flush.zero_()
        op.run(vin, vout)
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        op.run(vin, vout)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    tw = []
    for _ in range(steps):          # the reference's protocol: back-to-back, inputs unchanged, L2-warm (util.h:141-151)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        op.run(vin, vout)
        b.record()
        torch.cuda.synchronize()
        tw.append(a.elapsed_time(b))
    op.close()
    ms = float(np.mean(ts))
    peak, _ = measured_peaks()
    bm = bytes_min(m, nnz, k)
    return {"workload": f"{shape}_k{k}", "num_v": m, "nnz": nnz, "K": k, "ms_per_step": round(ms, 5),
            "ms_per_step_l2_warm": round(float(np.mean(tw)), 5),
            "gflops": round(2.0 * nnz * k / ms / 1e6, 1), "hbm_gbs_bytes_min": round(bm / ms / 1e6, 1),
            "roofline_frac": round(bm / ms / 1e6 / peak, 4), "gather_gbs": round(bytes_gather(m, nnz, k) / ms / 1e6, 1),
            "l2": "flushed between iterations"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import hpc_b200 as H

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: whatever NCCL prints while it initialises ("NCCL version ..." is a
        # bare printf at NCCL_DEBUG=VERSION) is sent to stderr by pointing fd 1 at fd 2 for the duration
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    shape, k = WORKLOADS[args.workload]
    H.set_host_threads(max(1, len(os.sched_getaffinity(0)) // world))   # the generator's share of the host cores
    ptr, idx = H.gen_named_graph(shape, SEED)
    m, nnz = len(ptr) - 1, len(idx)
    deg = np.diff(ptr)

    # nnz-balanced contiguous row partition; B replicated
    bounds = H.partition_rows(ptr, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    lptr = H.rebase_ptr(ptr, r0, r1)
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    lm, lnnz = r1 - r0, e1 - e0

    d_ptr = torch.from_numpy(lptr).to(dev)
    d_idx = torch.from_numpy(idx[e0:e1].copy() if world > 1 else idx).to(dev)
    val_full = torch.empty(nnz, dtype=torch.float32, device=dev)
    H.fill_normal(val_full, SEED, 1)
    d_val = val_full[e0:e1].clone() if world > 1 else val_full
    del val_full
    vin = torch.empty(m * k, dtype=torch.float32, device=dev)
    H.fill_normal(vin, SEED, 2)
    vout = torch.empty(max(1, lm * k), dtype=torch.float32, device=dev)
    g = H.CSR(lm, lnnz, d_ptr, d_idx, d_val)
    # The operator reads B rows by global column id, so it is built over num_v = local rows but
    # gathers from the full replicated B.
    op = H.SpMMB200(g, k, b_rows=m, **{o.split('=')[0]: int(o.split('=')[1]) for o in args.opt})
    t0 = time.perf_counter()
    op.preprocess(vin, vout)
    prep_s = time.perf_counter() - t0
    info = op.plan_info()

    working_set = bytes_min(lm, lnnz, k, b_rows=m)
    flush = None
    if working_set < 2 * L2_BYTES:
        flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        if flush is not None:
            flush.zero_()
        op.run(vin, vout)
    barrier()

    # ---- timed region: K steps, device time by CUDA events on the launching stream ----------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        barrier()
        wall0 = time.perf_counter()
        for a, b in evs:
            if flush is not None:
                flush.zero_()          # L2 flush between iterations, outside the event pair
            a.record()
            op.run(vin, vout)
            b.record()
        barrier()
        wall = time.perf_counter() - wall0
    step_ms = np.asarray([a.elapsed_time(b) for a, b in evs])
    t_ms = torch.tensor([float(step_ms.mean()), float(step_ms.min())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step, ms_min = float(t_ms[0]), float(t_ms[1])
    launches = op.launches_per_run * args.steps

    # ---- the kernel's own launch duration (events inside the C ABI, around the launch) --------
    kms = []
    for _ in range(max(3, min(10, args.steps))):
        if flush is not None:
            flush.zero_()
        kms.append(op.run_profiled(vin, vout))
    kernel_ms = float(np.mean(kms))

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -------------
    h_in = torch.empty(m * k, dtype=torch.float32).pin_memory()
    h_in.copy_(vin.cpu())
    h_out = torch.empty(max(1, lm * k), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        op.run_host(h_in, h_out)
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        op.run_host(h_in, h_out)       # H2D(B) -> kernels -> D2H(C) -> stream sync
    barrier()
    e2e_ms = (time.perf_counter() - te) / e2e_steps * 1e3
    te_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(te_t[0])
    checksum = float(h_out[: lm * k].double().sum()) if lm * k else 0.0

    if rank == 0:
        flops = 2.0 * nnz * k
        peak, peak_src = measured_peaks()
        total_bytes = bytes_min(lm, lnnz, k, b_rows=m)
        n_launch = max(1, op.launches_per_run)
        line = {
            "metric": "spmm_gflops", "value": round(flops / ms_per_step / 1e6, 2), "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5),
            "ms_per_step_min": round(ms_min, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": args.workload, "graph": shape, "num_v": m, "nnz": nnz, "K": k,
                "max_row_nnz": int(deg.max()), "mean_row_nnz": round(float(deg.mean()), 2),
                "p50_row_nnz": int(np.percentile(deg, 50)), "p99_row_nnz": int(np.percentile(deg, 99)),
                "empty_rows": int((deg == 0).sum()),
                "partition": f"rows by nnz over {world} rank(s), B replicated, no collective",
                "l2": "L2 flushed (256 MiB write) between iterations" if flush is not None
                      else "inputs larger than L2 (col/val + B re-streamed every step)",
                "plan": {kk: info[kk] for kk in ("seg_len", "kslice", "n_slices", "lanes", "vec", "n_col_blocks", "n_light", "n_heavy", "n_seg")},
                "preprocess_s": round(prep_s, 4),
            },
            "hbm_gbs_bytes_min": round(total_bytes / ms_per_step / 1e6, 1),
            "gather_gbs": round(bytes_gather(lm, lnnz, k) / ms_per_step / 1e6, 1),
            "roofline": {
                "bound": "hbm", "kernel": "spmm_kernel", "launches_per_step": n_launch,
                "achieved": round(total_bytes / kernel_ms / 1e6, 1), "peak": peak, "unit": "GB/s",
                "frac": round(total_bytes / kernel_ms / 1e6 / peak, 4),
                "traffic": NCU_TRAFFIC.get(args.workload) if world == 1 else None,
                "l2_hit_pct_ncu": NCU_L2_HIT_PCT.get(args.workload) if world == 1 else None,
                "peak_source": peak_src, "kernel_ms": round(kernel_ms / n_launch, 5),
                "algorithmic_bytes": int(total_bytes // n_launch),
                "gather_gbs": round(bytes_gather(lm, lnnz, k) / kernel_ms / 1e6, 1),
                "binding": binding_bound(args.workload, bytes_gather(lm, lnnz, k) / kernel_ms / 1e6, kernel_ms / n_launch, peak, world),
                "note": "per launch = per column-block pass (equal shares of the step); rank 0's partition; algorithmic bytes = "
                        "ptr+col+val+B once+C once (SURVEY.md 8d). The binding bound is the L2->SM gather of B rows (or HBM when B "
                        "is far larger than L2), not compulsory bytes: see DESIGN.md section 3 and profiles/r01_sweep.md",
            },
            "e2e": {"value": round(flops / e2e_ms / 1e6, 2), "unit": "GFLOP/s", "ms_per_step": round(e2e_ms, 4),
                    "h2d_bytes_per_step": 4 * m * k, "d2h_bytes_per_step": 4 * lm * k, "steps": e2e_steps,
                    "call": "spmm_b200_run_host (pinned B in, C out; CSR + plan resident, as in the reference harness)",
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "wall_s_timed_region": round(wall, 4),
        }
        if world == 1 and not args.no_also:
            # the other single-GPU BASELINE configs, kernel-only, so that one line covers K=32 and K=256
            line["also"] = [quick_measure(H, torch, sh, kk, dev) for sh, kk in (("arxiv", 32), ("arxiv", 256))
                            if f"{sh}_k{kk}" != args.workload]
        if world == 1 and not args.no_cpu_baseline:
            gf, desc, cores, _, _ = cpu_sample(ptr, idx, k, args.cpu_seconds)
            line["cpu_baseline"] = {"value": round(gf, 3), "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    op.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of spmm_kernel, per launch, from the `ncu --set full` captures
# summarised in profiles/r01_ncu_summary.md (1 GPU). reddit_k256: mean of the 5 column-block passes of one step.
NCU_TRAFFIC = {
    "reddit_k256": int(777.4e6),                  # mean of the 5 column-block passes of one step (prof_reddit_final)
    "products_k256": int(67.02e9 + 2.58e9),       # prof_products_k256_final
    "reddit_k32": int(1.176e9 + 34.6e6),          # prof_reddit_k32_final
    "arxiv_k32": int(33.74e6 + 0.18e6),           # prof_arxiv_k32_r01d
    "arxiv_k256": int(599.96e6 + 119.85e6),       # prof_arxiv_k256_r01d
}


# lts__t_sector_hit_rate.pct of the same captures: the L2 hit rate, essentially that of the B-row gathers
NCU_L2_HIT_PCT = {"reddit_k256": 94.9, "products_k256": 35.6, "reddit_k32": 86.6, "arxiv_k32": 55.2, "arxiv_k256": 32.0}

# What actually binds each workload (ncu, profiles/r01_ncu_summary.md). "l2_fabric": bytes gathered L2 -> SM per second
# against the fabric rate at which ncu shows lts2xbar 100 % busy (19.58 TB/s at 84.6 % => 23.1 TB/s). "hbm_traffic": the
# DRAM bytes ncu measured per launch, moved in the live kernel time, against the measured HBM peak.
BINDING = {"reddit_k256": "l2_fabric", "reddit_k32": "l2_fabric", "products_k256": "hbm_traffic",
           "arxiv_k256": "hbm_traffic", "arxiv_k32": "latency", "c0_k32": "latency"}
L2_FABRIC_PEAK_GBS = 23100.0


def binding_bound(workload, gather_gbs, kernel_ms_per_launch, peak_hbm, world):
    kind = BINDING.get(workload, "latency")
    if kind == "l2_fabric":
        return {"bound": "l2_fabric", "achieved": round(gather_gbs, 1), "peak": L2_FABRIC_PEAK_GBS, "unit": "GB/s",
                "frac": round(gather_gbs / L2_FABRIC_PEAK_GBS, 4), "source": "ncu lts2xbar_cycles_active (profiles/r01_ncu_summary.md)"}
    if kind == "hbm_traffic" and world == 1 and workload in NCU_TRAFFIC:
        a = NCU_TRAFFIC[workload] / kernel_ms_per_launch / 1e6
        return {"bound": "hbm_traffic", "achieved": round(a, 1), "peak": peak_hbm, "unit": "GB/s", "frac": round(a / peak_hbm, 4),
                "source": "ncu dram bytes per launch / live kernel time"}
    return {"bound": kind}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="reddit_k256", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the kernel-only numbers of the other configs")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (tuning runs only)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
