"""Exploration script (not part of the product or the tests): parity spot-checks and a timing
sweep over plan options on one B200. Usage: python tools/gpu_sweep.py [quick|full]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import hpc_b200 as H
from oracle import cpu as O

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3, cold=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def setup(name, K, seed=123):
    ptr, idx = H.gen_named_graph(name, seed)
    M, nnz = len(ptr) - 1, len(idx)
    d_ptr = torch.from_numpy(ptr).to(dev)
    d_idx = torch.from_numpy(idx).to(dev)
    val = torch.empty(nnz, dtype=torch.float32, device=dev)
    H.fill_normal(val, seed, 1)
    vin = torch.empty(M * K, dtype=torch.float32, device=dev)
    H.fill_normal(vin, seed, 2)
    vout = torch.full((M * K,), float("nan"), dtype=torch.float32, device=dev)
    g = H.CSR(M, nnz, d_ptr, d_idx, val)
    return ptr, idx, g, vin, vout


def parity(name, K, **opts):
    ptr, idx, g, vin, vout = setup(name, K)
    M = g.num_v
    op = H.SpMMB200(g, K, **opts)
    op.preprocess(vin, vout)
    op.run(vin, vout)
    torch.cuda.synchronize()
    got = vout.cpu().numpy().reshape(M, K)
    val = g.val.cpu().numpy()
    b = vin.cpu().numpy()
    # inputs equal the oracle's generator?
    same_in = np.array_equal(val, O.fill_normal(len(val), 123, 1)) and np.array_equal(b, O.fill_normal(len(b), 123, 2))
    ref = O.spmm_f32(ptr, idx, val, b, K)
    info = op.plan_info()
    heavy = sorted(op.heavy_row_set())
    hs = set(heavy)
    light = np.asarray([r for r in range(M) if r not in hs], np.int64)
    exact_light = np.array_equal(got[light].view(np.int32), ref[light].view(np.int32))
    ab = O.spmm_abssum(ptr, idx, val, b, K)
    err = np.abs(got.astype(np.float64) - ref) / np.maximum(ab, 1e-30)
    print(json.dumps({"parity": name, "K": K, "opts": opts, "inputs_match_oracle": bool(same_in),
                      "light_bit_exact": bool(exact_light), "n_heavy": int(len(heavy)),
                      "max_err_over_abssum": float(err.max()), "nan": int(np.isnan(got).sum()),
                      "plan": {k: info[k] for k in ("seg_len", "kslice", "lanes", "vec", "n_seg")}}), flush=True)


def bench(name, K, optlist, iters=10):
    ptr, idx, g, vin, vout = setup(name, K)
    M, nnz = g.num_v, g.num_e
    bytes_min = 4 * (M + 1) + 8 * nnz + 8 * M * K
    bytes_gather = 4 * (M + 1) + 8 * nnz + 4 * nnz * K + 4 * M * K
    for opts in optlist:
        op = H.SpMMB200(g, K, **opts)
        t0 = time.time()
        op.preprocess(vin, vout)
        tp = time.time() - t0
        mean_c, min_c = timeit(lambda: op.run(vin, vout), iters=iters, cold=True)
        mean_w, min_w = timeit(lambda: op.run(vin, vout), iters=iters, cold=False)
        info = op.plan_info()
        print(json.dumps({"bench": name, "K": K, "opts": opts, "ms_cold": round(mean_c, 4), "ms_cold_min": round(min_c, 4),
                          "ms_warm": round(mean_w, 4), "gflops": round(2 * nnz * K / mean_c / 1e6, 1),
                          "min_GBs": round(bytes_min / mean_c / 1e6, 1), "gather_GBs": round(bytes_gather / mean_c / 1e6, 1),
                          "prep_s": round(tp, 3), "seg_len": info["seg_len"], "kslice": info["kslice"],
                          "n_seg": info["n_seg"], "launches": op.launches_per_run}), flush=True)
        op.close()


print(torch.cuda.get_device_name(0), flush=True)
if mode in ("quick", "full"):
    parity("c0", 32)
    parity("arxiv", 256)
    for g, K in (("arxiv", 32), ("arxiv", 256), ("collab", 32), ("youtube", 32), ("reddit", 32), ("reddit", 64), ("reddit", 128), ("reddit", 256), ("products", 256), ("products", 32)):
        bench(g, K, [{}], iters=10 if g in ("arxiv", "collab", "youtube") else 5)
