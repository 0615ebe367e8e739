#!/usr/bin/env python
"""Lean ncu target: one 32^4 problem, a handful of Dslash launches (no solver, no linalg set-up kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from quda_b200 import dslash as D  # noqa: E402

pname, recon = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("single", 12)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
nsrc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
P = bench.make_device_problem([32, 32, 32, 32], bench.PREC_BYTES[pname], recon)
st = torch.cuda.current_stream().cuda_stream
tile = [int(v) for v in os.environ["PROF_TILE"].split()] if os.environ.get("PROF_TILE") else None
if nsrc > 1:
    srcs = [P["in"]] + [bench.new_spinor(P, seed=77 + i) for i in range(nsrc - 1)]
    dsts = [P["out"]] + [bench.new_spinor(P, seed=None) for i in range(nsrc - 1)]
# rotate through 4 (input, output) pairs as bench.py does, so that the captured launch writes its output back to HBM
pairs = [(P["in"], P["out"])] + [(bench.new_spinor(P, seed=501 + i), bench.new_spinor(P, seed=None)) for i in range(3)]
for k in range(n):
    if nsrc > 1:
        D.ApplyWilson(dsts, srcs, P["U"], 0.0, None, 0, 0, stream=st, tile=tile)
    else:
        D.ApplyWilson(pairs[k % 4][1], pairs[k % 4][0], P["U"], 0.0, None, 0, 0, stream=st, tile=tile)
torch.cuda.synchronize()
print("prof_target done", pname, recon)
