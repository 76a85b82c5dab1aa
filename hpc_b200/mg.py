"""Thin caller of the C ABI's multi-GPU driver (spmm_b200_mg_*, hpc_b200/csrc/multi.cu): one process, N devices.

No reference counterpart (single GPU). Host numpy arrays in, host tensors in/out for run_host;
the device-resident calls expose each device's buffers as raw pointers (`device_buffers`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


class MultiGpuSpMM:
    def __init__(self, ptr: np.ndarray, idx: np.ndarray, val: np.ndarray, feat_in: int, devices=None, n_devices: int = 0,
                 **options):
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.num_v, self.num_e, self.feat_in = len(self.ptr) - 1, len(self.idx), int(feat_in)
        devs = list(devices) if devices is not None else list(range(n_devices or 1))
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        check(lib.spmm_b200_mg_create(self.ptr.ctypes.data_as(C.c_void_p), self.idx.ctypes.data_as(C.c_void_p),
                                      self.val.ctypes.data_as(C.c_void_p), self.num_v, self.num_e, self.feat_in,
                                      len(devs), C.cast(arr, C.c_void_p), C.byref(h)))
        self._h = h
        self.devices = devs
        for k, v in options.items():
            check(lib.spmm_b200_mg_set_option(self._h, k.encode(), int(v)))

    def info(self) -> dict:
        n, p2p = C.c_int(0), C.c_int(0)
        bounds = np.empty(len(self.devices) + 1, np.int32)
        check(lib.spmm_b200_mg_info(self._h, C.byref(n), bounds.ctypes.data_as(C.c_void_p), C.byref(p2p)))
        return {"n_devices": n.value, "bounds": bounds, "peer_access": bool(p2p.value)}

    def preprocess(self) -> None:
        check(lib.spmm_b200_mg_preprocess(self._h))

    def device_buffers(self, g: int) -> dict:
        b, c, full, op = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib.spmm_b200_mg_device_buffers(self._h, g, C.byref(b), C.byref(c), C.byref(full), C.byref(op)))
        return {"b": b.value or 0, "c_block": c.value or 0, "c_full": full.value or 0, "op": op.value or 0}

    def run_host(self, h_vin, h_vout) -> None:
        """h_vin: B (num_v * feat_in), h_vout: C, host float32 tensors / arrays (pinned for overlap)."""
        check(lib.spmm_b200_mg_run_host(self._h, C.c_void_p(_addr(h_vin)), C.c_void_p(_addr(h_vout))))

    def run(self) -> None:
        check(lib.spmm_b200_mg_run(self._h))

    def sync(self) -> None:
        check(lib.spmm_b200_mg_sync(self._h))

    def set_fused(self, on: bool) -> None:
        check(lib.spmm_b200_mg_set_fused(self._h, 1 if on else 0))

    def allgather(self) -> None:
        check(lib.spmm_b200_mg_allgather(self._h))

    def swap(self) -> None:
        check(lib.spmm_b200_mg_swap(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.spmm_b200_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _addr(t) -> int:
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
