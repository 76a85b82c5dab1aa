"""Timing of the transposed operator (dB = A^T dC) beside the forward one, same graph, K=256."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H
for shape, K in (("arxiv", 256), ("reddit", 256), ("products", 256)):
    ptr, idx = H.gen_named_graph(shape)
    M, nnz = len(ptr) - 1, len(idx)
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1))
    x = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
    y = torch.empty(M * K, device="cuda")
    fwd = H.SpMMB200(g, K)
    fwd.preprocess(x, y)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    bwd = fwd.transposed()
    b.record(); torch.cuda.synchronize()
    t_build = a.elapsed_time(b)
    bwd.preprocess(y, x)
    def t(op, i, o):
        for _ in range(3): op.run(i, o)
        return float(np.mean([op.run_profiled(i, o) for _ in range(10)]))
    tf, tb = t(fwd, x, y), t(bwd, y, x)
    info = bwd.plan_info()
    print(json.dumps({"shape": shape, "K": K, "ms_forward": round(tf, 4), "ms_transposed": round(tb, 4), "ms_build_transposed_csr": round(t_build, 2),
                      "transposed_plan": {k: info[k] for k in ("n_col_blocks", "n_heavy", "n_seg", "reorder")}}), flush=True)
    bwd.close(); fwd.close()
