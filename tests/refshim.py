"""ctypes doorway to oracle/_ref/libspmm_ref.so — the UNMODIFIED reference kernels of liblaf/hpc
PA4 rebuilt for sm_100a by oracle/Makefile (test infrastructure only; needs a GPU to run)."""
import ctypes as C
import os

from conftest import ROOT

_PATH = os.path.join(ROOT, "oracle", "_ref", "libspmm_ref.so")
_STUDENT = os.path.join(ROOT, "oracle", "_ref", "libspmm_student.so")


def available():
    return os.path.exists(_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        P, I = C.c_void_p, C.c_int
        _lib.ref_spmm_run.argtypes = [P, P, P, P, P, I, I, I]
        _lib.ref_spmm_run.restype = I
        _lib.ref_valid_float.argtypes = [P, P, I]
        _lib.ref_valid_float.restype = I
        _lib.ref_cusparse_run.argtypes = [P, P, P, P, P, I, I, I, I, C.POINTER(C.c_double)]
        _lib.ref_cusparse_run.restype = I
        _lib.ref_load_graph.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(I), C.POINTER(I), P, P, C.c_longlong,
                                        C.c_longlong]
        _lib.ref_load_graph.restype = I
    return _lib


def ref_spmm(d_ptr, d_idx, d_val, d_vin, d_vout, m, nnz, k):
    """SpMMRef::preprocess + run (PA4/handout/src/spmm_ref.cu:20-30) on torch CUDA tensors."""
    rc = lib().ref_spmm_run(d_ptr.data_ptr(), d_idx.data_ptr(), d_val.data_ptr(), d_vin.data_ptr(),
                            d_vout.data_ptr(), m, nnz, k)
    assert rc == 0, rc


def ref_cusparse(d_ptr, d_idx, d_val, d_vin, d_vout, m, nnz, k):
    sec = C.c_double(0)
    rc = lib().ref_cusparse_run(d_ptr.data_ptr(), d_idx.data_ptr(), d_val.data_ptr(), d_vin.data_ptr(),
                                d_vout.data_ptr(), m, nnz, k, 0, C.byref(sec))
    assert rc == 0, rc


def ref_valid(d_y, d_y2, n):
    return lib().ref_valid_float(d_y.data_ptr(), d_y2.data_ptr(), n)
