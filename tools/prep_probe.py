"""Where does preprocess time go? (exploration)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H
for shape, K in [("reddit", 32), ("reddit", 256), ("products", 256), ("arxiv", 32)]:
    ptr, idx = H.gen_named_graph(shape)
    M, nnz = len(ptr) - 1, len(idx)
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 1, 1))
    vin = torch.zeros(M * K, device="cuda"); vout = torch.empty(M * K, device="cuda")
    for rep in range(3):
        op = H.SpMMB200(g, K)
        torch.cuda.synchronize(); t = time.perf_counter()
        op.preprocess(vin, vout)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        info = op.plan_info()
        print(shape, K, "rep", rep, "preprocess_s", round(dt, 4), {k: info[k] for k in ("n_light", "n_ltask", "n_seg", "n_col_blocks", "light_steps")}, flush=True)
        op.close()
