"""Debug probe for the persistent launch: each case in its own subprocess under a hard timeout (a hung kernel must not
eat the GPU budget). Usage: python tools/persist_probe.py"""
import subprocess
import sys

CASES = [
    ("c0", 256, dict(persistent=1)),
    ("c0", 256, dict(persistent=1, col_blocks=3, seg_len=32, reorder=0, row_groups=4)),
    ("c0", 32, dict(persistent=1, col_blocks=3, seg_len=16)),
    ("arxiv", 256, dict(persistent=1, col_blocks=4, reorder=0)),
    ("arxiv", 256, dict(persistent=1, col_blocks=4, reorder=1)),
    ("reddit", 256, dict()),
]
CHILD = r'''
import sys, json, time
import numpy as np, torch
sys.path.insert(0, ".")
import hpc_b200 as H
shape, K, opts = sys.argv[1], int(sys.argv[2]), json.loads(sys.argv[3])
ptr, idx = H.gen_named_graph(shape)
M, nnz = len(ptr) - 1, len(idx)
g = H.CSR(M, nnz, torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda(), H.fill_normal(torch.empty(nnz, device="cuda"), 123, 1))
vin = H.fill_normal(torch.empty(M * K, device="cuda"), 123, 2)
ref_op = H.SpMMB200(g, K, **{k: v for k, v in opts.items() if k not in ("persistent", "row_groups")}, persistent=0)
want = torch.empty(M * K, device="cuda")
ref_op.preprocess(vin, want); ref_op.run(vin, want); torch.cuda.synchronize()
op = H.SpMMB200(g, K, **opts)
got = torch.full((M * K,), float("nan"), device="cuda")
op.preprocess(vin, got)
info = op.plan_info()
print("plan", {k: info[k] for k in ("persistent", "n_row_groups", "n_tickets", "n_col_blocks", "reorder", "lanes")}, flush=True)
for rep in range(6):
    got.fill_(float("nan"))
    t = time.perf_counter()
    op.run(vin, got); torch.cuda.synchronize()
    print("run", rep, round((time.perf_counter() - t) * 1e3, 3), "ms equal:", bool(torch.equal(got, want)), "launches", op.launches_per_run,
          "counters nonzero:", int(np.count_nonzero(op.plan_arrays()["counters"])), flush=True)
'''
import json
for shape, K, opts in CASES:
    print("==", shape, K, opts, flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, shape, str(K), json.dumps(opts)], capture_output=True, text=True, timeout=90)
        print(r.stdout[-1500:], r.stderr[-800:], flush=True)
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT", (e.stdout or b"")[-1500:], flush=True)
