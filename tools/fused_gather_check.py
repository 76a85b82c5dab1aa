"""Multi-GPU check + timing of the stacked-layer epilogue (torchrun, one rank per GPU):
two stacked SpMM layers, (a) SpMM + NCCL all-gather-v of C, (b) SpMM whose epilogue stores C rows into every
rank's peer-mapped buffer (no collective). Outputs must be bit-identical. Usage:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29600 tools/fused_gather_check.py [shape] [K]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpc_b200 as H
from hpc_b200.dist import ShardedSpMM

shape = sys.argv[1] if len(sys.argv) > 1 else "products"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

ptr, idx = H.gen_named_graph(shape)
M, nnz = len(ptr) - 1, len(idx)
val = H.fill_normal(torch.empty(nnz, device=dev), 123, 1).cpu().numpy()
b0 = H.fill_normal(torch.empty(M * K, device=dev), 123, 2)


def barrier():
    dist.barrier()
    torch.cuda.synchronize()


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


# (a) NCCL path
sh = ShardedSpMM(ptr, idx, val, K, device=dev)
local_c = torch.empty(max(1, sh.local_rows * K), device=dev)
full1 = torch.empty(M * K, device=dev)
full2 = torch.empty(M * K, device=dev)
sh.preprocess(b0, local_c)


def two_layers_nccl():
    sh.run(b0, local_c)
    sh.allgather(local_c, full1)
    sh.run(full1, local_c)
    sh.allgather(local_c, full2)


two_layers_nccl()
barrier()
ref1, ref2 = full1.clone(), full2.clone()
t_spmm = timed(lambda: sh.run(b0, local_c))
t_nccl = timed(two_layers_nccl)

# (b) fused epilogue
shf = ShardedSpMM(ptr, idx, val, K, device=dev)
bufs = shf.enable_fused_gather(n_buffers=2, use_multicast=os.environ.get("SPMM_NO_MC") is None)
local_f = torch.empty(max(1, shf.local_rows * K), device=dev)
shf.preprocess(b0, local_f)


def two_layers_fused():
    c1 = shf.run_fused(b0, local_f, buffer=0)
    return shf.run_fused(c1, local_f, buffer=1)


for b in bufs:
    b.fill_(float("nan"))
barrier()
two_layers_fused()
barrier()
ok1 = bool(torch.equal(bufs[0], ref1))
ok2 = bool(torch.equal(bufs[1], ref2))
t_fused = timed(two_layers_fused)
oks = torch.tensor([int(ok1), int(ok2)], device=dev)
dist.all_reduce(oks, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"shape": shape, "K": K, "world": world, "multicast": bool(shf._use_mc), "layer1_bit_equal": bool(oks[0]),
                      "layer2_bit_equal": bool(oks[1]), "ms_spmm_only": round(t_spmm, 4), "ms_two_layers_nccl": round(t_nccl, 4),
                      "ms_two_layers_fused": round(t_fused, 4), "allgather_bytes_per_layer": 4 * M * K}), flush=True)
sh.close()
shf.close()
dist.barrier()
dist.destroy_process_group()
