"""Experiment: can the kernel's row epilogue stream C straight into pinned host memory (zero-copy stores over PCIe)
fast enough to replace the D2H copy of the host-buffer call?  Usage: python tools/zero_copy_probe.py [shape] [K]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import hpc_b200 as H

shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ptr, idx = H.gen_named_graph(shape)
M, nnz = len(ptr) - 1, len(idx)
g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1))
vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
vout = torch.empty(M * K, device="cuda")
h_in = vin.cpu().pin_memory()
h_out = torch.empty(M * K).pin_memory()


def wall(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


op = H.SpMMB200(g, K)
op.preprocess(vin, vout)
res = {"shape": shape, "K": K, "bytes_c": 4 * M * K}
res["ms_run"] = wall(lambda: op.run(vin, vout))
res["ms_run_plus_d2h"] = wall(lambda: (op.run(vin, vout), h_out.copy_(vout, non_blocking=True)))
res["ms_d2h_alone"] = wall(lambda: h_out.copy_(vout, non_blocking=True))
res["ms_run_host"] = wall(lambda: op.run_host(h_in, h_out))
ref = vout.cpu()
op.close()
op = H.SpMMB200(g, K)
op.set_gather([h_out.data_ptr()], 0)          # pinned host memory is device-addressable under UVA
op.preprocess(vin, vout)
h_out.fill_(float("nan"))
res["ms_run_zero_copy_store"] = wall(lambda: op.run(vin, vout))
res["zero_copy_bit_equal"] = bool(torch.equal(h_out, ref))
res["zero_copy_gbs"] = res["bytes_c"] / res["ms_run_zero_copy_store"] / 1e6
print(json.dumps(res), flush=True)
op.close()
