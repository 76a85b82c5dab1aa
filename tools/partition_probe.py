"""Exploration: one rank's row block of an N-way partition, timed on one GPU with different numbers of column blocks."""
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
for parts in (8, 4, 2):
    bounds = H.partition_rows(ptr, parts)
    r0, r1 = int(bounds[0]), int(bounds[1])
    lptr = H.rebase_ptr(ptr, r0, r1)
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    g = H.CSR(r1 - r0, e1 - e0, torch.from_numpy(lptr).cuda(), d_idx[e0:e1].clone(), val[e0:e1].clone())
    vout = torch.empty((r1 - r0) * K, device="cuda")
    for nb in (0, 1, 2, 3, 4, 5, 8):
        op = H.SpMMB200(g, K, b_rows=M, col_blocks=nb)
        op.preprocess(vin, vout)
        for _ in range(3): op.run(vin, vout)
        ts = [op.run_profiled(vin, vout) for _ in range(10)]
        print(json.dumps({"parts": parts, "rows": r1 - r0, "nnz": e1 - e0, "col_blocks": nb, "effective": op.plan_info()["n_col_blocks"], "ms": round(float(np.mean(ts)), 4)}), flush=True)
        op.close()
