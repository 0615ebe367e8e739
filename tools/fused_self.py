#!/usr/bin/env python
"""Single-GPU timing / profiling target for the partitioned schedules: the rank is its own neighbour ("self" exchange: pack
writes into its own ghost slabs and raises its own arrival flags), so the full pack | interior | boundary machinery runs in one
process and can be put under ncu.  usage: fused_self.py "0 1 1 1" [steps] [prec recon]   (B200_HALO_SCHEDULE=streams for
the round-1 two-stream schedule)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from quda_b200 import comm, dirac as DR  # noqa: E402

mask = [int(v) for v in sys.argv[1].split()] if len(sys.argv) > 1 else [0, 0, 0, 1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
pname = sys.argv[3] if len(sys.argv) > 3 else "single"
recon = int(sys.argv[4]) if len(sys.argv) > 4 else 12
prec = bench.PREC_BYTES[pname]
X = [32, 32, 32, 32]
P = bench.make_device_problem(X, prec, recon)
st = torch.cuda.current_stream().cuda_stream
pairs = [(P["in"], P["out"])] + [(bench.new_spinor(P, seed=501 + i), bench.new_spinor(P, seed=None)) for i in range(3)]
ex = comm.HaloExchange(comm.ProcessGrid((1, 1, 1, 1), 0), X, prec, mode="self", self_dims=mask)
cs = ex.comm_struct()
op = DR.Dirac("wilson", P["U"], 0.0, comm=cs, stream=st)
for i in range(10):
    op.Dslash(pairs[i % 4][1], pairs[i % 4][0], 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    op.Dslash(pairs[i % 4][1], pairs[i % 4][0], 0)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / steps * 1e3
print(json.dumps({"mask": mask, "schedule": os.environ.get("B200_HALO_SCHEDULE", "fused"), "prec": pname, "recon": recon,
                  "us_per_dslash": us, "timed_out": bool(ex.timed_out())}))
