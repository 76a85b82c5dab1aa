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

    def __init__(self, ptr: np.ndarray, world: int):
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.world = int(world)
        self.num_v = len(self.ptr) - 1
        self.bounds = partition_rows(self.ptr, self.world)

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
    g = CSR(len(local_ptr) - 1, len(local_idx),
            torch.from_numpy(local_ptr).to(device), torch.from_numpy(np.ascontiguousarray(local_idx)).to(device),
            local_val.to(device) if isinstance(local_val, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_val)).to(device))
    return SpMMB200(g, feat, b_rows=b_rows, **options)


class ShardedSpMM:
    """SpMM over a row partition. Every rank holds the full B (`vin`, num_v x feat) and produces
    its own row block of C; `allgather` assembles the full C on every rank (e.g. as the next
    layer's B)."""

    def __init__(self, ptr, idx, val, feat: int, group=None, device="cuda", op_factory=None, **options):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.feat = int(feat)
        self.part = RowPartition(ptr, self.world)
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
        if backend == "nccl":
            # one NCCL all-gather of equal (padded) blocks into a staging buffer, then one
            # device-side compaction per peer block
            mx = max(elems)
            if getattr(self, "_stage", None) is None or self._stage.numel() < mx * self.world:
                self._stage = torch.empty(mx * self.world, dtype=full.dtype, device=full.device)
                self._pad = torch.zeros(mx, dtype=full.dtype, device=full.device)
            self._pad[: elems[self.rank]].copy_(vout_local[: elems[self.rank]])
            dist.all_gather_into_tensor(self._stage, self._pad, group=self.group)
            for g in range(self.world):
                full[offs[g]: offs[g + 1]].copy_(self._stage[g * mx: g * mx + elems[g]])
        else:
            # gloo (CPU tests): one broadcast per block, straight into place
            works = []
            for g in range(self.world):
                view = full[offs[g]: offs[g + 1]]
                if g == self.rank:
                    view.copy_(vout_local[: elems[g]])
                if elems[g]:
                    works.append(dist.broadcast(view, src=dist.get_global_rank(self.group, g) if self.group else g,
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

    def close(self):
        if hasattr(self.op, "close"):
            self.op.close()
