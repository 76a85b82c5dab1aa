#!/usr/bin/env python
"""run_all.sh restated (PA4/handout/script/run_all.sh:3-11): the 13 datasets, one unit_tests run each, output
tee'd to a log whose `time = ... (double)` lines PA4/workspace/plot.py:13-27 parses. The datasets are the
synthetic stand-ins of hpc_b200/graph.py (same rows / nnz / max row nnz). Usage: run_all.py [--len 32] [names...]"""
import argparse
import datetime
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hpc_b200.graph import GRAPH_SHAPES, RUN_ALL_DATASETS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--len", type=int, default=32)
ap.add_argument("--log", default=None)
ap.add_argument("names", nargs="*", default=list(RUN_ALL_DATASETS))
args = ap.parse_args()
log = args.log or f"output_{datetime.datetime.now():%H_%M_%S_%m_%d}.log"
print("Log saved to", log)
failed = 0
with open(log, "a") as f:
    for name in args.names:
        print(name, flush=True)
        spec = ",".join(str(x) for x in GRAPH_SHAPES[name])
        r = subprocess.run([os.path.join(HERE, "unit_tests"), "--gen", spec, "--dataset", name, "--len", str(args.len)],
                           capture_output=True, text=True)
        out = r.stderr + r.stdout
        sys.stdout.write(out)
        f.write(out)
        failed += r.returncode != 0
sys.exit(1 if failed else 0)
