#!/usr/bin/env python
"""Dev tool (GPU): time the interior kernel over launch tilings for several (precision, recon) pairs in ONE process,
for the library selected by $B200_LIB.  Output: gpurun_out/tune_<tag>.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from quda_b200 import dslash as D, fields as F  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
cfgs = [("single", 12), ("double", 18), ("half", 8), ("half", 12), ("single", 8), ("single", 18), ("double", 12)]
if len(sys.argv) > 2:
    cfgs = [tuple(c.split(":")) for c in sys.argv[2].split(",")]
    cfgs = [(p, int(r)) for p, r in cfgs]
X = [32, 32, 32, 32]
TILES = [(16, 1, 1, 1), (16, 2, 1, 1), (16, 2, 2, 1), (16, 2, 1, 2), (16, 4, 1, 1), (16, 4, 2, 1), (16, 4, 1, 2),
         (16, 8, 1, 1), (16, 2, 2, 2), (16, 1, 2, 1), (16, 1, 1, 2), (8, 8, 1, 1), (8, 4, 2, 1)]
stream = torch.cuda.current_stream().cuda_stream
res = {}
t_all = time.time()
for pname, recon in cfgs:
    prec = bench.PREC_BYTES[pname]
    t0 = time.time()
    P = bench.make_device_problem(X, prec, recon)
    Vh = F.volume_cb(X)
    bmin = D.min_bytes_per_site(prec, recon)
    rows = []
    for t in TILES:
        for _ in range(3):
            D.ApplyWilson(P["out"], P["in"], P["U"], 0.0, None, 0, 0, tile=t, stream=stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 30
        for _ in range(n):
            D.ApplyWilson(P["out"], P["in"], P["U"], 0.0, None, 0, 0, tile=t, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        rows.append({"tile": t, "us": us, "gbs": bmin * Vh / us * 1e-3, "gflops": 1320 * Vh / us * 1e-3})
    rows.sort(key=lambda r: r["us"])
    res[f"{pname}-r{recon}"] = rows
    print(f"tune[{tag}] {pname} r{recon}: best {rows[0]} worst {rows[-1]['us']:.1f}us (setup+sweep {time.time() - t0:.1f}s)", flush=True)
    del P
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tune_{tag}.json"), "w"))
print(f"tune[{tag}] total {time.time() - t_all:.1f}s")
