"""Multi-GPU driver: A's rows partitioned by nnz over the ranks of a torch.distributed group,
B replicated, C row blocks all-gathered only when a stacked layer needs the full output.

No reference counterpart (the reference is single-GPU: SURVEY.md §8e; the only trace is the
commented-out `// extern ncclComm_t* comms;` at PA4/handout/include/util.h:30). One process per
GPU; torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .graph import partition_rows, rebase_ptr


class RowPartition:
    """Contiguous row blocks balanced by nnz: rank g owns rows [bounds[g], bounds[g+1])."""

    def __init__(self, ptr: np.ndarray, world: int, row_cost: int = 0):
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.world = int(world)
        self.num_v = len(self.ptr) - 1
        self.bounds = partition_rows(self.ptr, self.world, row_cost)

    def rows(self, rank: int):
        return int(self.bounds[rank]), int(self.bounds[rank + 1])

    def nnz_range(self, rank: int):
        r0, r1 = self.rows(rank)
        return int(self.ptr[r0]), int(self.ptr[r1])

    def local_ptr(self, rank: int) -> np.ndarray:
        return rebase_ptr(self.ptr, *self.rows(rank))

    def block_elems(self, feat: int):
        return [int(self.bounds[g + 1] - self.bounds[g]) * feat for g in range(self.world)]


def _default_op_factory(local_ptr, local_idx, local_val, feat, b_rows, device, options):
    from .spmm import CSR, SpMMB200
    if isinstance(local_val, torch.Tensor):
        d_val = local_val.to(device)
        if d_val.untyped_storage().nbytes() > d_val.numel() * d_val.element_size():
            d_val = d_val.clone()   # a slice of the whole graph's values: keep only this block alive
    else:
        d_val = torch.from_numpy(np.ascontiguousarray(local_val)).to(device)
    g = CSR(len(local_ptr) - 1, len(local_idx),
            torch.from_numpy(local_ptr).to(device), torch.from_numpy(np.ascontiguousarray(local_idx)).to(device), d_val)
    return SpMMB200(g, feat, b_rows=b_rows, **options)


class ShardedSpMM:
    """SpMM over a row partition. Every rank holds the full B (`vin`, num_v x feat) and produces
    its own row block of C; `allgather` assembles the full C on every rank (e.g. as the next
    layer's B)."""

    def __init__(self, ptr, idx, val, feat: int, group=None, device="cuda", op_factory=None, row_cost: int = 0, **options):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.feat = int(feat)
        self.part = RowPartition(ptr, self.world, row_cost)
        self.num_v = self.part.num_v
        self.row_begin, self.row_end = self.part.rows(self.rank)
        e0, e1 = self.part.nnz_range(self.rank)
        self.local_rows = self.row_end - self.row_begin
        factory = op_factory or _default_op_factory
        self.op = factory(self.part.local_ptr(self.rank), idx[e0:e1], val[e0:e1], self.feat, self.num_v, device, options)

    def preprocess(self, vin, vout_local) -> None:
        self.op.preprocess(vin, vout_local)

    def run(self, vin, vout_local) -> None:
        """vout_local[local_rows * feat] = A[row_begin:row_end, :] @ vin. No communication."""
        self.op.run(vin, vout_local)

    def allgather(self, vout_local: torch.Tensor, full: torch.Tensor) -> torch.Tensor:
        """full[num_v * feat] <- every rank's row block, in row order. Blocks differ in size (the
        partition balances nnz, not rows), so this is an all-gather-v."""
        elems = self.part.block_elems(self.feat)
        offs = np.concatenate([[0], np.cumsum(elems)])
        if self.world == 1:
            full[: elems[0]].copy_(vout_local[: elems[0]])
            return full
        backend = dist.get_backend(self.group)
        views = [full[offs[g]: offs[g + 1]] for g in range(self.world)]
        if backend == "nccl":
            # all-gather-v straight into place: with unequal block sizes ProcessGroupNCCL issues one coalesced group of
            # ncclBroadcast calls (root g -> views[g]); no padding, no staging buffer, no compaction copies
            dist.all_gather(views, vout_local[: elems[self.rank]], group=self.group)
        else:
            # gloo (CPU tests): one broadcast per block, straight into place
            works = []
            for g in range(self.world):
                if g == self.rank:
                    views[g].copy_(vout_local[: elems[g]])
                if elems[g]:
                    works.append(dist.broadcast(views[g], src=dist.get_global_rank(self.group, g) if self.group else g,
                                                group=self.group, async_op=True))
            for w in works:
                w.wait()
        return full

    # ---- fused epilogue: the kernel itself delivers C rows to every rank (NVLink peer / multicast stores) ----

    def enable_fused_gather(self, n_buffers: int = 2, use_multicast: bool = True):
        """Allocates `n_buffers` symmetric (peer-mapped) full-size C buffers through torch's symmetric memory
        (plumbing only) and points the operator's epilogue at them. Call before preprocess. Layers alternate
        between the buffers so a layer's input is never overwritten by its own output."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group or dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self._sym_bufs, self._sym_hdls = [], []
        for _ in range(n_buffers):
            t = symm_mem.empty(self.num_v * self.feat, dtype=torch.float32, device=dev)
            self._sym_bufs.append(t)
            self._sym_hdls.append(symm_mem.rendezvous(t, group))
        self._use_mc = bool(use_multicast and getattr(self._sym_hdls[0], "has_multicast_support", lambda *a: False) is not None
                            and self._sym_hdls[0].multicast_ptr)
        self._select_gather(0)
        return self._sym_bufs

    def _select_gather(self, i: int):
        h = self._sym_hdls[i]
        mc = h.multicast_ptr if self._use_mc else 0
        self.op.set_gather([int(p) for p in h.buffer_ptrs], self.row_begin, multicast=mc)
        self._sym_cur = i

    def run_fused(self, vin, vout_local, buffer: int = 0) -> torch.Tensor:
        """SpMM whose epilogue also writes this rank's rows into every rank's symmetric buffer `buffer`; after
        the cross-rank barrier the returned tensor holds the full C on every rank. No collective call."""
        if self._sym_cur != buffer:
            self._select_gather(buffer)
        self.op.run(vin, vout_local)
        self._sym_hdls[buffer].barrier()
        return self._sym_bufs[buffer]

    # ---- sharded host I/O: B replicated over NVLink instead of N uploads over PCIe ------------------------------

    def enable_sharded_host_io(self, use_multicast: bool = True) -> torch.Tensor:
        """Allocates this rank's full copy of B and the barrier flag words as symmetric memory (torch's
        `_symmetric_memory`: allocation + address exchange only) and hands every rank's mapping to the operator
        (spmm_b200_set_replicate). Rank g then uploads the g-th of `world` pieces of every ~48 MB chunk of B."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group or dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self._rep_b = symm_mem.empty(self.num_v * self.feat, dtype=torch.float32, device=dev)
        self._rep_flags = symm_mem.empty(max(64, 2 * self.world), dtype=torch.int32, device=dev)
        self._rep_flags.zero_()
        hb = symm_mem.rendezvous(self._rep_b, group)
        hf = symm_mem.rendezvous(self._rep_flags, group)
        torch.cuda.synchronize()
        dist.barrier(group)   # every rank's flags are zero before anybody signals
        mc = (hb.multicast_ptr or 0) if use_multicast else 0
        self._rep_handles = (hb, hf)
        self.rep_multicast = bool(mc)
        self.op.set_replicate(self.world, self.rank, hb.buffer_ptrs, mc, hf.buffer_ptrs)
        return self._rep_b

    def run_host_sharded(self, h_vin: torch.Tensor, h_vout_local: torch.Tensor) -> int:
        """h_vin: the whole B on the host (this rank reads only its 1/world share of every chunk); h_vout_local: its
        block of C. Collective: every rank calls it once per step. Returns the bytes this rank uploaded."""
        return self.op.run_host_sharded(h_vin, h_vout_local)

    def close(self):
        if hasattr(self.op, "close"):
            self.op.close()
