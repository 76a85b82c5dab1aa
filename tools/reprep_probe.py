"""Exploration: what preprocess / destroy cost on a handle that is re-planned, and over create -> preprocess -> destroy cycles
(SPMM_B200_PREP_TRACE=1 prints the phases; SPMM_B200_POOL=0 selects cudaMalloc / cudaFree for the plan arrays)."""
import json, os, sys, time
import torch
sys.path.insert(0, ".")
import hpc_b200 as H
pool = os.environ.get("SPMM_B200_POOL", "1")
for shape, K in [("reddit", 256), ("products", 256), ("arxiv", 256)]:
    ptr, idx = H.gen_named_graph(shape)
    M, nnz = len(ptr) - 1, len(idx)
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 1, 1))
    vin = torch.zeros(M * K, device="cuda"); vout = torch.empty(M * K, device="cuda")
    op = H.SpMMB200(g, K)
    same = []
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        op.preprocess(vin, vout)
        torch.cuda.synchronize(); same.append(round(time.perf_counter() - t, 4))
        op.run(vin, vout); torch.cuda.synchronize()
    t = time.perf_counter(); op.close(); torch.cuda.synchronize()
    destroy = round(time.perf_counter() - t, 4)
    cycles = []
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        op = H.SpMMB200(g, K); op.preprocess(vin, vout); torch.cuda.synchronize()
        t1 = time.perf_counter()
        op.run(vin, vout); torch.cuda.synchronize()
        t2 = time.perf_counter(); op.close(); torch.cuda.synchronize()
        cycles.append([round(t1 - t, 4), round(time.perf_counter() - t2, 4)])
    t = time.perf_counter(); H.trim_memory(); trim = round(time.perf_counter() - t, 4)
    print(json.dumps({"pool": pool, "shape": shape, "K": K, "preprocess_s_same_handle": same, "destroy_s": destroy,
                      "create_preprocess_s__destroy_s": cycles, "trim_s": trim}), flush=True)
