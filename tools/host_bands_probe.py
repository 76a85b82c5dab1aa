"""run_host with equal column blocks vs the host_bands layout (large last block), reddit K=256, one GPU."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H
ptr, idx = H.gen_named_graph("reddit")
M, nnz, K = len(ptr) - 1, len(idx), 256
g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1))
vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
vout = torch.empty(M * K, device="cuda")
h_in, h_out = vin.cpu().pin_memory(), torch.empty(M * K).pin_memory()
ref = None
for opts in ({}, {"host_bands": 1}, {"host_bands": 30}, {"host_bands": 40}, {"host_bands": 60}, {"host_bands": 70},
             {"host_bands": 40, "col_blocks": 4}, {"host_bands": 60, "col_blocks": 4}, {"host_bands": 1, "zero_copy": 0}):
    op = H.SpMMB200(g, K, **opts)
    op.preprocess(vin, vout)
    for _ in range(3): op.run_host(h_in, h_out)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): op.run_host(h_in, h_out)
    e2e = (time.perf_counter() - t) / 10 * 1e3
    dev = float(np.mean([op.run_profiled(vin, vout) for _ in range(5)]))
    if ref is None: ref = h_out.clone()
    bands = [(op.plan_info(b)["col_begin"], op.plan_info(b)["col_end"]) for b in range(op.plan_info()["n_col_blocks"])]
    print(json.dumps({**opts, "e2e_ms": round(e2e, 3), "device_ms": round(dev, 3), "bit_equal": bool(torch.equal(h_out, ref)), "bands": bands}), flush=True)
    op.close()
