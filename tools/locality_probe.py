"""Exploration: how much of the L2 locality a graph offers does the kernel realise? products-shaped graph with
all-local / all-global / mixed columns, natural vs bucketed row order. Prints ms and effective gather GB/s."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H

def run(ptr, idx, K, **opts):
    M, nnz = len(ptr) - 1, len(idx)
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 1, 1))
    vin = H.fill_normal(torch.empty(M * K, device="cuda"), 1, 2)
    vout = torch.empty(M * K, device="cuda")
    op = H.SpMMB200(g, K, **opts); op.preprocess(vin, vout)
    for _ in range(3): op.run(vin, vout)
    ts = [op.run_profiled(vin, vout) for _ in range(5)]
    op.close()
    return float(np.mean(ts))

M, nnz, mx = 2449029, 123718280, 17481
for local_ppm, window in [(1000000, 8192), (500000, 8192), (0, 8192), (1000000, 65536)]:
    ptr, idx = H.gen_graph(M, nnz, mx, 2, 20000, local_ppm, window, seed=123)
    for opts in ({}, {"reorder": 0, "block": 32}):
        ms = run(ptr, idx, 256, **opts)
        print(json.dumps({"local_ppm": local_ppm, "window": window, "opts": opts, "ms": round(ms, 3),
                          "gather_TBs": round(nnz * 1024 / ms / 1e9, 2)}), flush=True)
