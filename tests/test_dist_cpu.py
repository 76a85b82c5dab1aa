"""world_size-2 (and 3) gloo tests of the multi-GPU driver's host logic on CPU.

The row-block SpMM itself needs a GPU, so these tests inject an oracle-backed stand-in for the
local operator (tests may use the oracle as the checker's compute); what is exercised is the
product's partitioning, rebasing, slicing and the all-gather-v assembly of hpc_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hpc_b200 as H
from hpc_b200.dist import RowPartition, ShardedSpMM
from oracle import cpu as O


class OracleOp:
    """CPU stand-in with SpMMB200's preprocess/run surface."""

    def __init__(self, lptr, lidx, lval, feat, b_rows, device, options):
        self.ptr, self.idx, self.val, self.feat = lptr, np.asarray(lidx), np.asarray(lval), feat

    def preprocess(self, vin, vout):
        pass

    def run(self, vin, vout):
        out = O.spmm_f32(self.ptr, self.idx, self.val, vin.numpy(), self.feat)
        vout[: out.size].copy_(torch.from_numpy(out.ravel()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ptr, idx = H.gen_named_graph("c0")
        M, nnz = len(ptr) - 1, len(idx)
        val = O.fill_normal(nnz, 123, 1)
        b = torch.from_numpy(O.fill_normal(M * K, 123, 2))
        sh = ShardedSpMM(ptr, idx, val, K, device="cpu", op_factory=OracleOp)
        local = torch.full((max(1, sh.local_rows * K),), float("nan"))
        sh.preprocess(b, local)
        sh.run(b, local)
        full = torch.full((M * K,), float("nan"))
        sh.allgather(local, full)
        want = O.spmm_f32(ptr, idx, val, b.numpy(), K).ravel()
        ok = np.array_equal(full.numpy().view(np.int32), want.view(np.int32))
        # two stacked layers: layer 2 consumes the gathered C as its B
        local2 = torch.empty_like(local)
        sh.run(full, local2)
        full2 = torch.empty(M * K)
        sh.allgather(local2, full2)
        want2 = O.spmm_f32(ptr, idx, val, want, K).ravel()
        ok2 = np.array_equal(full2.numpy().view(np.int32), want2.view(np.int32))
        q.put((rank, bool(ok), bool(ok2), sh.row_begin, sh.row_end))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_spmm_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    # blocks tile the rows in rank order
    assert res[0][3] == 0 and res[-1][4] == 4096
    assert all(res[i][4] == res[i + 1][3] for i in range(world - 1))


def test_row_partition_helpers():
    ptr, _ = H.gen_named_graph("c0")
    part = RowPartition(ptr, 4)
    assert sum(part.block_elems(8)) == 4096 * 8
    tot = 0
    for g in range(4):
        e0, e1 = part.nnz_range(g)
        lp = part.local_ptr(g)
        assert lp[0] == 0 and lp[-1] == e1 - e0
        tot += e1 - e0
    assert tot == ptr[-1]


def test_row_partition_cost_balanced():
    """RowPartition with the plan's per-row overhead (row_cost): the weighted rule of spmm_b200_partition_rows_weighted,
    restated in oracle/plan_oracle.py; blocks still tile the rows and the slices line up."""
    from oracle import plan_oracle as P
    ptr, _ = H.gen_named_graph("arxiv")
    for world, rc in ((4, 5), (8, 1)):
        part = RowPartition(ptr, world, row_cost=rc)
        assert np.array_equal(part.bounds, P.partition_rows(ptr, world, rc))
        assert part.bounds[0] == 0 and part.bounds[-1] == len(ptr) - 1
        tot = 0
        for g in range(world):
            e0, e1 = part.nnz_range(g)
            lp = part.local_ptr(g)
            assert lp[0] == 0 and lp[-1] == e1 - e0
            tot += e1 - e0
        assert tot == ptr[-1]
