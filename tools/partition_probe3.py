"""parts = 8 block of the reddit shape: persistent launch with shorter tasks / segments (do the waits go away?)."""
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
for parts in (8, 16):
    bounds = H.partition_rows(ptr, parts)
    r0, r1 = int(bounds[0]), int(bounds[1])
    lptr = H.rebase_ptr(ptr, r0, r1)
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    g = H.CSR(r1 - r0, e1 - e0, torch.from_numpy(lptr).cuda(), d_idx[e0:e1].clone(), val[e0:e1].clone())
    vout = torch.empty((r1 - r0) * K, device="cuda")
    for opts in ({"persistent": 0}, {"persistent": 1}, {"persistent": 1, "seg_len": 128}, {"persistent": 1, "seg_len": 64},
                 {"persistent": 0, "seg_len": 128}, {"persistent": 1, "seg_len": 128, "light_steps": 32}, {"persistent": 1, "col_blocks": 3},
                 {"persistent": 0, "col_blocks": 3}, {"persistent": 1, "col_blocks": 4}, {"persistent": 0, "col_blocks": 4}):
        op = H.SpMMB200(g, K, b_rows=M, **opts)
        op.preprocess(vin, vout)
        for _ in range(5): op.run(vin, vout)
        ts = []
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(20): op.run(vin, vout)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 20)
        print(json.dumps({"parts": parts, **opts, "ms": [round(x, 4) for x in ts]}), flush=True)
        op.close()
