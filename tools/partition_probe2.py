"""Persistent single launch vs one launch per column block, interleaved repetitions, one rank's block of an N-way partition."""
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
for parts in (1, 2, 3, 4, 6, 8):
    bounds = H.partition_rows(ptr, parts)
    r0, r1 = int(bounds[0]), int(bounds[1])
    lptr = H.rebase_ptr(ptr, r0, r1)
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    g = H.CSR(r1 - r0, e1 - e0, torch.from_numpy(lptr).cuda(), d_idx[e0:e1].clone(), val[e0:e1].clone())
    vout = torch.empty((r1 - r0) * K, device="cuda")
    ops = {}
    for name, opts in (("multi", {"persistent": 0, "split_streams": 0}), ("persistent", {"persistent": 1}),
                       ("split", {"persistent": 0, "split_streams": 1})):
        op = H.SpMMB200(g, K, b_rows=M, **opts)
        op.preprocess(vin, vout)
        for _ in range(5): op.run(vin, vout)
        ops[name] = op
    res = {"multi": [], "persistent": [], "split": []}
    for rep in range(4):
        for name, op in ops.items():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(20): op.run(vin, vout)
            b.record(); torch.cuda.synchronize()
            res[name].append(a.elapsed_time(b) / 20)
    outs = []
    for name, op in ops.items():
        o = torch.full_like(vout, float("nan"))
        op.run(vin, o)
        torch.cuda.synchronize()
        outs.append(o)
    eq = all(torch.equal(outs[0], o) for o in outs[1:])
    info = ops["persistent"].plan_info()
    bands = [ops["multi"].plan_info(b)["n_utask"] for b in range(info["n_col_blocks"])]
    print(json.dumps({"parts": parts, "band_tasks": bands, "waves_shortest_band": round(min(bands) / info["resident_warps"], 1),
                      "ms_multi": [round(x, 4) for x in res["multi"]], "ms_persistent": [round(x, 4) for x in res["persistent"]],
                      "ms_split": [round(x, 4) for x in res["split"]],
                      "ratio_persistent": round(float(np.mean(res["persistent"]) / np.mean(res["multi"])), 4),
                      "ratio_split": round(float(np.mean(res["split"]) / np.mean(res["multi"])), 4),
                      "bit_equal": bool(eq)}), flush=True)
    for op in ops.values(): op.close()
