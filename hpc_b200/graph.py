"""Graphs: synthetic generator, the reference's on-disk format, nnz-balanced row partition.

load_graph / write_graph follow PA4/handout/src/data.cu:3-66. The shapes below stand in for the
OGB / DGL graphs the reference reads from ~/PA4/data (PA4/handout/script/run_all.sh:3-11):
rows and nnz from the public dataset cards, max row nnz from PA4/workspace/phase_2.log.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

from .shapes import GRAPH_SHAPES, RUN_ALL_DATASETS  # noqa: E402,F401


def _ip(a):
    return a.ctypes.data_as(C.c_void_p)


def set_host_threads(n: int) -> None:
    """OpenMP threads for the generator (torchrun sets OMP_NUM_THREADS=1 in its workers)."""
    check(lib.spmm_b200_set_host_threads(int(n)))


def gen_degrees(num_v, nnz, max_deg, tail_k=2, zero_ppm=0, seed=123) -> np.ndarray:
    deg = np.empty(num_v, dtype=np.int32)
    check(lib.spmm_b200_gen_degrees(num_v, nnz, max_deg, tail_k, zero_ppm, seed, _ip(deg)))
    return deg


def gen_graph(num_v, nnz, max_deg, tail_k=2, zero_ppm=0, local_ppm=500000, window=1024, seed=123):
    """-> (ptr int32[num_v+1], idx int32[nnz]); columns ascending and unique within a row."""
    ptr = np.empty(num_v + 1, dtype=np.int32)
    idx = np.empty(nnz, dtype=np.int32)
    check(lib.spmm_b200_gen_graph(num_v, nnz, max_deg, tail_k, zero_ppm, local_ppm, window, seed,
                                  _ip(ptr), _ip(idx)))
    return ptr, idx


def gen_named_graph(name: str, seed: int = 123):
    return gen_graph(*GRAPH_SHAPES[name], seed=seed)


def load_graph(datadir: str, dset: str):
    nv, ne = C.c_int(0), C.c_int(0)
    check(lib.spmm_b200_load_graph(datadir.encode(), dset.encode(), C.byref(nv), C.byref(ne), None, None))
    ptr = np.empty(nv.value + 1, dtype=np.int32)
    idx = np.empty(ne.value, dtype=np.int32)
    check(lib.spmm_b200_load_graph(datadir.encode(), dset.encode(), C.byref(nv), C.byref(ne), _ip(ptr), _ip(idx)))
    return nv.value, ne.value, ptr, idx


def write_graph(datadir: str, dset: str, ptr: np.ndarray, idx: np.ndarray, text: bool = False) -> None:
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    check(lib.spmm_b200_write_graph(datadir.encode(), dset.encode(), len(ptr) - 1, len(idx), _ip(ptr), _ip(idx),
                                    1 if text else 0))


def partition_rows(ptr: np.ndarray, parts: int, row_cost: int = 0) -> np.ndarray:
    """bounds int32[parts+1]: contiguous row blocks balanced by nnz, or by nnz + row_cost per row
    (the plan's per-row overhead; plan_row_cost) when row_cost > 0."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    bounds = np.empty(parts + 1, dtype=np.int32)
    if row_cost:
        check(lib.spmm_b200_partition_rows_weighted(_ip(ptr), len(ptr) - 1, parts, int(row_cost), _ip(bounds)))
    else:
        check(lib.spmm_b200_partition_rows(_ip(ptr), len(ptr) - 1, parts, _ip(bounds)))
    return bounds


def plan_row_cost(num_v: int, nnz: int, feat: int, b_rows: int = 0) -> int:
    """Header entries per row of the automatic plan (= its column blocks): the row_cost to partition with."""
    return int(lib.spmm_b200_plan_row_cost(int(num_v), int(nnz), int(b_rows), int(feat)))


def rebase_ptr(ptr: np.ndarray, row_begin: int, row_end: int) -> np.ndarray:
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    out = np.empty(row_end - row_begin + 1, dtype=np.int32)
    check(lib.spmm_b200_rebase_ptr(_ip(ptr), row_begin, row_end, _ip(out)))
    return out
