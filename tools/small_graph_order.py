"""arxiv-shaped graph, K=256 and K=32: natural vs bucketed row order (now that natural order is planned in 16 row groups)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H
ptr, idx = H.gen_named_graph("arxiv")
M, nnz = len(ptr) - 1, len(idx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K in (256, 32):
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1))
    vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
    vout = torch.empty(M * K, device="cuda")
    for opts in ({}, {"reorder": 0}, {"reorder": 0, "row_groups": 1}, {"reorder": 0, "row_groups": 64}, {"reorder": 1, "row_groups": 16},
                 {"reorder": 1, "row_groups": 64}):
        op = H.SpMMB200(g, K, **opts)
        op.preprocess(vin, vout)
        for _ in range(3): op.run(vin, vout)
        cold = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); op.run(vin, vout); b.record(); torch.cuda.synchronize()
            cold.append(a.elapsed_time(b))
        print(json.dumps({"K": K, **opts, "ms_cold": round(float(np.mean(cold)), 4), "groups": op.plan_info()["n_row_groups"]}), flush=True)
        op.close()
