"""arxiv-shaped graph: task size (light_steps) and segment length sweep, L2 flushed between runs (as bench.py does) and warm."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H
ptr, idx = H.gen_named_graph("arxiv")
M, nnz = len(ptr) - 1, len(idx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for K, steps_list in ((32, (0, 8, 16, 24, 32, 48, 64, 96)), (256, (0, 32, 64, 96, 128, 192, 256))):
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1))
    vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
    vout = torch.empty(M * K, device="cuda")
    for seg_len in (0, 64, 512):
        for steps in steps_list:
            op = H.SpMMB200(g, K, light_steps=steps, seg_len=seg_len)
            op.preprocess(vin, vout)
            for _ in range(3): op.run(vin, vout)
            cold, warm = [], []
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); op.run(vin, vout); b.record(); torch.cuda.synchronize()
                cold.append(a.elapsed_time(b))
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); op.run(vin, vout); b.record(); torch.cuda.synchronize()
                warm.append(a.elapsed_time(b))
            info = op.plan_info()
            print(json.dumps({"K": K, "seg_len": info["seg_len"], "light_steps": info["light_steps"], "tasks": info["n_utask"],
                              "waves": round(info["n_utask"] / info["resident_warps"], 2), "ms_cold": round(float(np.mean(cold)), 4),
                              "ms_warm": round(float(np.mean(warm)), 4)}), flush=True)
            op.close()
