"""Multi-rank GPU checks, run under torch.distributed.run by tests/test_multi_gpu.py (one rank per GPU):
  * sharded host I/O (spmm_b200_run_host_sharded): every rank uploads 1/N of B, the ranks replicate it over NVLink
    (multicast and peer-store variants), output bit-equal to the device-resident run and to the 1-GPU result;
  * stacked layers: the kernel's own epilogue (multimem.st and peer stores) against the NCCL all-gather-v, bit-equal.
Prints one JSON line on rank 0; exit code != 0 on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpc_b200 as H  # noqa: E402
from hpc_b200.dist import ShardedSpMM  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "c0"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
opts = {o.split("=")[0]: int(o.split("=")[1]) for o in sys.argv[3:]}
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

ptr, idx = H.gen_named_graph(shape)
M, nnz = len(ptr) - 1, len(idx)
val = H.fill_normal(torch.empty(nnz, device=dev), 123, 1)
b0 = H.fill_normal(torch.empty(M * K, device=dev), 123, 2)

# the single-GPU result of two stacked layers (every rank computes it: it is the checker here)
g = H.CSR(M, nnz, torch.from_numpy(ptr).to(dev), torch.from_numpy(idx).to(dev), val)
one = H.SpMMB200(g, K, **opts)
want1, want2 = torch.empty(M * K, device=dev), torch.empty(M * K, device=dev)
one.preprocess(b0, want1)
one.run(b0, want1)
one.run(want1, want2)
torch.cuda.synchronize()
one.close()

res = {"shape": shape, "K": K, "world": world, "opts": opts}
ok = True
for use_mc in (True, False):
    tag = "mc" if use_mc else "peer"
    sh = ShardedSpMM(ptr, idx, val, K, device=dev, **opts)
    r0, r1 = sh.row_begin, sh.row_end
    local_c = torch.full((max(1, sh.local_rows * K),), float("nan"), device=dev)
    sh.preprocess(b0, local_c)
    sh.run(b0, local_c)
    torch.cuda.synchronize()
    ok_dev = torch.equal(local_c[: sh.local_rows * K], want1[r0 * K: r1 * K])
    # sharded host I/O
    sh.enable_sharded_host_io(use_multicast=use_mc)
    res[f"{tag}_replicate_multicast"] = bool(sh.rep_multicast)
    h_in = b0.cpu().pin_memory()
    outs = []
    for pinned in (True, False):
        h_out = torch.full((max(1, sh.local_rows * K),), float("nan"))
        if pinned:
            h_out = h_out.pin_memory()
        for _ in range(3):                       # repeated calls: the phase-0 barrier orders pushes against earlier passes
            sh.run_host_sharded(h_in, h_out)
        outs.append(torch.equal(h_out[: sh.local_rows * K], want1[r0 * K: r1 * K].cpu()))
    # stacked layers: NCCL all-gather-v vs the fused epilogue
    full1, full2 = torch.empty(M * K, device=dev), torch.empty(M * K, device=dev)
    sh.run(b0, local_c)
    sh.allgather(local_c, full1)
    sh.run(full1, local_c)
    sh.allgather(local_c, full2)
    torch.cuda.synchronize()
    ok_nccl = torch.equal(full1, want1) and torch.equal(full2, want2)
    shf = ShardedSpMM(ptr, idx, val, K, device=dev, **opts)
    bufs = shf.enable_fused_gather(n_buffers=2, use_multicast=use_mc)
    res[f"{tag}_gather_multicast"] = bool(shf._use_mc)
    local_f = torch.empty(max(1, shf.local_rows * K), device=dev)
    shf.preprocess(b0, local_f)
    for bb in bufs:
        bb.fill_(float("nan"))
    dist.barrier()
    torch.cuda.synchronize()
    c1 = shf.run_fused(b0, local_f, buffer=0)
    c2 = shf.run_fused(c1, local_f, buffer=1)
    torch.cuda.synchronize()
    ok_fused = torch.equal(c1, want1) and torch.equal(c2, want2)
    flags = torch.tensor([int(ok_dev), int(outs[0]), int(outs[1]), int(ok_nccl), int(ok_fused)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    for name, v in zip(("device_run", "host_sharded_pinned", "host_sharded_pageable", "nccl_allgather", "fused_epilogue"), flags.tolist()):
        res[f"{tag}_{name}"] = bool(v)
        ok &= bool(v)
    sh.close()
    shf.close()
    dist.barrier()
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
