"""ctypes access to oracle/liboracle.so (spmm_oracle.c) — TEST INFRASTRUCTURE ONLY.

Restates PA4/handout/src/spmm_ref.cu:3-17 (SpMM), src/valid.cu:3-25 (validators) and
PA4/workspace/src/spmm_opt.cu:43-54 (student task split) on the CPU; see the C file's header
for the parity status. Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build() -> None:
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


def _load():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "spmm_oracle.c")):
        build()
    lib = C.CDLL(_SO)
    P, I, LL, U64, F = C.c_void_p, C.c_int, C.c_longlong, C.c_uint64, C.c_float
    lib.oracle_spmm_literal.argtypes = [P, P, P, P, P, I, I, I]
    lib.oracle_spmm_literal.restype = None
    lib.oracle_spmm_f32.argtypes = [P, P, P, P, P, I, I, I, I, I]
    lib.oracle_spmm_f32.restype = None
    lib.oracle_spmm_t_f32.argtypes = [P, P, P, P, P, P, I, I, I, I]
    lib.oracle_spmm_t_f32.restype = None
    lib.oracle_spmm_f64.argtypes = [P, P, P, P, P, I, I, I, I]
    lib.oracle_spmm_f64.restype = None
    lib.oracle_spmm_abssum.argtypes = [P, P, P, P, P, I, I, I]
    lib.oracle_spmm_abssum.restype = None
    lib.oracle_validate_float.argtypes = [P, P, LL]
    lib.oracle_validate_float.restype = LL
    lib.oracle_validate_int.argtypes = [P, P, LL]
    lib.oracle_validate_int.restype = LL
    lib.oracle_fill_normal.argtypes = [P, LL, U64, U64, F, F]
    lib.oracle_fill_normal.restype = None
    lib.oracle_student_tasks.argtypes = [P, I, I, P]
    lib.oracle_student_tasks.restype = LL
    lib.oracle_num_threads.argtypes = []
    lib.oracle_num_threads.restype = I
    return lib


_lib = _load()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _csr(ptr, idx, val, vin):
    return (np.ascontiguousarray(ptr, np.int32), np.ascontiguousarray(idx, np.int32),
            np.ascontiguousarray(val, np.float32), np.ascontiguousarray(vin, np.float32))


def spmm_literal(ptr, idx, val, vin, feat, ftz=False):
    """spmm_ref.cu:3-17 in its literal loop order (slow; small cases)."""
    ptr, idx, val, vin = _csr(ptr, idx, val, vin)
    m = len(ptr) - 1
    out = np.empty(m * feat, np.float32)
    _lib.oracle_spmm_literal(_p(ptr), _p(idx), _p(val), _p(vin), _p(out), m, feat, int(ftz))
    return out.reshape(m, feat)


def spmm_f32(ptr, idx, val, vin, feat, row_begin=0, row_end=None, ftz=False, nthreads=0, out=None):
    """Same chains, loops interchanged and OpenMP over rows; rows [row_begin, row_end) only."""
    ptr, idx, val, vin = _csr(ptr, idx, val, vin)
    m = len(ptr) - 1
    row_end = m if row_end is None else row_end
    if out is None:
        out = np.zeros(m * feat, np.float32)
    _lib.oracle_spmm_f32(_p(ptr), _p(idx), _p(val), _p(vin), _p(out), feat, row_begin, row_end, int(ftz), nthreads)
    return out.reshape(m, feat)


def spmm_t_f32(ptr, idx, val, dc, feat, b_rows=None, ftz=False, with_abs=False):
    """dB = A^T dC: one in-order FMA chain per output element over its column's nonzeros in CSR storage order
    (no reference counterpart: parity unpinned). -> dB[b_rows, feat] (and the fp64 sum of |terms| with with_abs)."""
    ptr, idx, val, dc = _csr(ptr, idx, val, dc)
    m = len(ptr) - 1
    b_rows = m if b_rows is None else b_rows
    out = np.empty(b_rows * feat, np.float32)
    ab = np.empty(b_rows * feat, np.float64) if with_abs else None
    _lib.oracle_spmm_t_f32(_p(ptr), _p(idx), _p(val), _p(dc), _p(out), _p(ab) if with_abs else None, m, b_rows, feat, int(ftz))
    out = out.reshape(b_rows, feat)
    return (out, ab.reshape(b_rows, feat)) if with_abs else out


def spmm_f64(ptr, idx, val, vin, feat):
    ptr, idx, val, vin = _csr(ptr, idx, val, vin)
    m = len(ptr) - 1
    out = np.zeros(m * feat, np.float64)
    _lib.oracle_spmm_f64(_p(ptr), _p(idx), _p(val), _p(vin), _p(out), feat, 0, m, 0)
    return out.reshape(m, feat)


def spmm_abssum(ptr, idx, val, vin, feat):
    ptr, idx, val, vin = _csr(ptr, idx, val, vin)
    m = len(ptr) - 1
    out = np.zeros(m * feat, np.float64)
    _lib.oracle_spmm_abssum(_p(ptr), _p(idx), _p(val), _p(vin), _p(out), feat, 0, m)
    return out.reshape(m, feat)


def validate_float(ref, ans):
    """valid.cu:3-13: count of |(ref - ans) / ref| > 1e-2 (first argument normalises)."""
    ref = np.ascontiguousarray(ref, np.float32).ravel()
    ans = np.ascontiguousarray(ans, np.float32).ravel()
    return int(_lib.oracle_validate_float(_p(ref), _p(ans), ref.size))


def validate_int(ref, ans):
    ref = np.ascontiguousarray(ref, np.int32).ravel()
    ans = np.ascontiguousarray(ans, np.int32).ravel()
    return int(_lib.oracle_validate_int(_p(ref), _p(ans), ref.size))


def fill_normal(n, seed, stream, mean=0.0, stddev=0.1):
    out = np.empty(n, np.float32)
    _lib.oracle_fill_normal(_p(out), n, seed, stream, mean, stddev)
    return out


def student_tasks(ptr, batch=256):
    ptr = np.ascontiguousarray(ptr, np.int32)
    n = _lib.oracle_student_tasks(_p(ptr), len(ptr) - 1, batch, None)
    out = np.empty(3 * n, np.int32)
    _lib.oracle_student_tasks(_p(ptr), len(ptr) - 1, batch, _p(out))
    return out.reshape(-1, 3)


def num_threads():
    return int(_lib.oracle_num_threads())
