"""Context numbers on the same B200 (off the product path): the reference's own
kernels rebuilt for sm_100a from /root/reference by oracle/Makefile — spmm_kernel_ref
(PA4/handout/src/spmm_ref.cu), cuSPARSE as the handout calls it (src/spmm_cusparse.cu) and the
student's SpMMOpt (PA4/workspace/src/spmm_opt.cu) — timed beside the engine, L2-warm
(the reference protocol, util.h:141-151) with CUDA events. Usage: python tools/compare_ref.py [shapes...]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpc_b200 as H  # noqa: E402
import refshim  # noqa: E402

student = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libspmm_student.so"))
student.student_spmm_run.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.POINTER(C.c_double)]
student.student_spmm_run.restype = C.c_int


def ev_time(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


cases = sys.argv[1:] or ["arxiv:32", "arxiv:256", "reddit:32", "reddit:256", "products:256"]
for case in cases:
    shape, K = case.split(":")
    K = int(K)
    ptr, idx = H.gen_named_graph(shape)
    M, nnz = len(ptr) - 1, len(idx)
    d_ptr, d_idx = torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda()
    val = H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1)
    vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
    vout = torch.zeros(M * K, device="cuda")
    vref = torch.zeros(M * K, device="cuda")
    g = H.CSR(M, nnz, d_ptr, d_idx, val)
    op = H.SpMMB200(g, K)
    op.preprocess(vin, vout)
    out = {"graph": shape, "K": K, "num_v": M, "nnz": nnz}
    out["engine_ms"] = ev_time(lambda: op.run(vin, vout), 10, 20)
    sec = C.c_double(0)
    rc = refshim.lib().ref_cusparse_run(d_ptr.data_ptr(), d_idx.data_ptr(), val.data_ptr(), vin.data_ptr(), vref.data_ptr(),
                                        M, nnz, K, 1, C.byref(sec))
    out["cusparse_ms"] = sec.value * 1e3 if rc == 0 else None
    out["cusparse_vs_engine_mismatch"] = H.valid(vout, vref, M * K)
    heavy = nnz * K > 4e9          # the thread-per-row kernel needs seconds per run on the big graphs
    iters = 1 if heavy else 5
    out["ref_kernel_ms"] = ev_time(lambda: refshim.ref_spmm(d_ptr, d_idx, val, vin, vref, M, nnz, K), 0 if heavy else 2, iters)
    out["ref_vs_engine_mismatch"] = H.valid(vout, vref, M * K)
    sec = C.c_double(0)
    vst = torch.zeros(M * K, device="cuda")
    rc = student.student_spmm_run(d_ptr.data_ptr(), d_idx.data_ptr(), val.data_ptr(), vin.data_ptr(), vst.data_ptr(), M, nnz, K, 1, C.byref(sec))
    out["student_ms"] = sec.value * 1e3 if rc == 0 else None
    for k in ("cusparse_ms", "ref_kernel_ms", "student_ms"):
        if out[k]:
            out[k.replace("_ms", "_over_engine")] = round(out[k] / out["engine_ms"], 2)
    print(json.dumps(out), flush=True)
    op.close()
