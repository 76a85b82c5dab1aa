"""Small cases for compute-sanitizer (one tool per gpurun call): every kernel path on tiny inputs."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import hpc_b200 as H

ptr, idx = H.gen_named_graph("c0")
M, nnz = len(ptr) - 1, len(idx)
for K, opts in [(32, {}), (32, {"seg_len": 16}), (256, {"col_blocks": 3, "seg_len": 32}), (100, {}), (30, {}), (64, {"reorder": 0, "light_steps": 7})]:
    g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 1, 1))
    vin = H.fill_normal(torch.empty(M * K, device="cuda"), 1, 2)
    vout = torch.empty(M * K, device="cuda")
    tgt = [torch.zeros(M * K, device="cuda")] if K % 4 == 0 else []
    op = H.SpMMB200(g, K, **opts)
    if tgt:
        op.set_gather(tgt, 0)
    op.preprocess(vin, vout)
    op.run(vin, vout)
    op.run(vin, vout)
    torch.cuda.synchronize()
    assert torch.isfinite(vout).all()
    assert H.valid(vout, vout, M * K) == 0
    op.close()
print("sanitize cases done")
