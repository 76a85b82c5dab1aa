"""Multi-GPU parity on real devices (needs >= 2 visible GPUs; the CPU-side logic is covered by tests/test_dist_cpu.py):
the NVLink paths — multimem.st / peer-store epilogue and the sharded host I/O — against the single-GPU result."""
import json
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import ROOT  # noqa: E402

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 CUDA devices", allow_module_level=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("shape,K,extra", [("c0", 32, []), ("arxiv", 256, []), ("c0", 256, ["col_blocks=3", "seg_len=32"])])
def test_nvlink_paths_match_single_gpu(shape, K, extra):
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mp_checks.py"), shape, str(K)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert all(v for k, v in line.items() if k.endswith(("device_run", "host_sharded_pinned", "host_sharded_pageable",
                                                          "nccl_allgather", "fused_epilogue"))), line
