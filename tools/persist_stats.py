"""Persistent launch statistics (needs the stats build: `make -C hpc_b200/csrc stats`, then
SPMM_B200_LIB=$PWD/hpc_b200/libspmm_b200_stats.so python tools/persist_stats.py):
polls of the row-group counters, blocked waits and their total clocks, per run, for one rank's block of an N-way partition."""
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
for parts in (8, 4):
    bounds = H.partition_rows(ptr, parts)
    r0, r1 = int(bounds[0]), int(bounds[1])
    lptr = H.rebase_ptr(ptr, r0, r1)
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    g = H.CSR(r1 - r0, e1 - e0, torch.from_numpy(lptr).cuda(), d_idx[e0:e1].clone(), val[e0:e1].clone())
    vout = torch.empty((r1 - r0) * K, device="cuda")
    for opts in ({"persistent": 1}, {"persistent": 1, "ticket_batch": 4}, {"persistent": 1, "ticket_batch": 2}, {"persistent": 1, "light_steps": 128}, {"persistent": 0}):
        op = H.SpMMB200(g, K, b_rows=M, **opts)
        op.preprocess(vin, vout)
        for _ in range(3): op.run(vin, vout)
        torch.cuda.synchronize()
        def stats():
            c = op.plan_arrays()["counters"].view(np.uint32)
            w = c[-32:]
            return w[8:14].view(np.uint64).astype(np.int64)
        if not opts["persistent"]:
            print(json.dumps({"parts": parts, **opts, "ms": round(op.run_profiled(vin, vout), 4),
                              "per_band_tasks": [op.plan_info(b)["n_utask"] for b in range(op.plan_info()["n_col_blocks"])]}), flush=True)
            op.close()
            continue
        s0 = stats()
        ms = op.run_profiled(vin, vout)
        s1 = stats() - s0
        info = op.plan_info()
        print(json.dumps({"parts": parts, **opts, "tickets": info["n_tickets"], "groups": info["n_row_groups"], "ms": round(ms, 4),
                          "polls": int(s1[0]), "blocked_waits": int(s1[1]), "wait_ms_summed_over_warps": round(float(s1[2]) / 1.9e6, 3)}), flush=True)
        op.close()
