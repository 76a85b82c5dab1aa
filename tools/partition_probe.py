"""Exploration: one rank's row block of an N-way partition, timed on one GPU — persistent single launch vs one launch
per column block, a few task sizes. Events around single runs (run_profiled) and around back-to-back loops."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H

shape, K = "reddit", 256
ptr, idx = H.gen_named_graph(shape)
M, nnz = len(ptr) - 1, len(idx)
val = H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1)
vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
d_idx = torch.from_numpy(idx).cuda()
for parts in (8, 4, 1):
    bounds = H.partition_rows(ptr, parts)
    r0, r1 = int(bounds[0]), int(bounds[1])
    lptr = H.rebase_ptr(ptr, r0, r1)
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    g = H.CSR(r1 - r0, e1 - e0, torch.from_numpy(lptr).cuda(), d_idx[e0:e1].clone(), val[e0:e1].clone())
    vout = torch.empty((r1 - r0) * K, device="cuda")
    for opts in ({"persistent": 0}, {"persistent": 1}, {"persistent": 1, "row_groups": 4}, {"persistent": 1, "row_groups": 64},
                 {"persistent": 0, "light_steps": 64}, {"persistent": 1, "light_steps": 64}, {"persistent": 1, "light_steps": 256}):
        op = H.SpMMB200(g, K, b_rows=M, **opts)
        op.preprocess(vin, vout)
        for _ in range(5): op.run(vin, vout)
        ts = [op.run_profiled(vin, vout) for _ in range(20)]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(50): op.run(vin, vout)
        b.record(); torch.cuda.synchronize()
        info = op.plan_info()
        print(json.dumps({"parts": parts, "nnz": e1 - e0, **opts, "launches": op.launches_per_run, "tickets": info["n_tickets"],
                          "ms_single": round(float(np.mean(ts)), 4), "ms_min": round(float(np.min(ts)), 4),
                          "ms_back_to_back": round(a.elapsed_time(b) / 50, 4)}), flush=True)
        op.close()
