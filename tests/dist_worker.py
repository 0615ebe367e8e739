"""Shared multi-rank worker: every rank owns one block of a global lattice, exchanges halos with its neighbours and
applies x + a D in; the result must match the oracle on the GLOBAL lattice.  mode "host": gloo + host twin (CPU tier);
mode "p2p"/"nccl": one GPU per rank, libquda_b200.so, NVLink peer writes / NCCL send-recv (GPU tier)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, grid_dims, Xl, prec, recon, q, mode="host", reps=3):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch
    import torch.distributed as dist
    if mode == "host":
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import oracle
    from common import CudaMem, HostMem, twin_backend
    from quda_b200 import comm, dslash as D, fields as F
    be = twin_backend() if mode == "host" else None
    mem = HostMem if mode == "host" else CudaMem
    grid = comm.ProcessGrid(grid_dims, rank)
    Xg = [Xl[d] * grid_dims[d] for d in range(4)]
    hp = 8 if prec == 8 else 4
    gauge = oracle.random_gauge(Xg, hp, seed=137)          # same global field on every rank (same seed)
    parity, dagger, kappa = 1, 0, 0.12195
    s = oracle.random_spinor(Xg, hp, seed=5)
    xs = oracle.random_spinor(Xg, hp, seed=6)
    ref = xs.astype(np.float64) - kappa * oracle.wil_dslash(gauge, s, Xg, parity, dagger).astype(np.float64)
    # local pieces; the gauge pad gets the backward neighbours' links
    gl = comm.local_slice(gauge, Xg, Xl, grid.coords, "gauge")
    ghost_from = []
    for d in range(4):
        c = list(grid.coords)
        c[d] = (c[d] - 1) % grid_dims[d]
        ghost_from.append(comm.local_slice(gauge, Xg, Xl, c, "gauge") if grid_dims[d] > 1 else None)
    gbuf, gmeta = F.gauge_to_native(gl, Xl, prec, recon, ghost_from=ghost_from)
    U = D.GaugeField(mem.put(gbuf), Xl, prec, recon, gmeta, t_boundary=-1, first_time_slice=grid.first_time_slice(),
                     last_time_slice=grid.last_time_slice())
    sl = comm.local_slice(s, Xg, Xl, grid.coords, ("spinor1", 1 - parity))
    xl = comm.local_slice(xs, Xg, Xl, grid.coords, ("spinor1", parity))
    din = D.ColorSpinorField(mem.put(F.spinor_to_native(sl, prec)), Xl, prec)
    dx = D.ColorSpinorField(mem.put(F.spinor_to_native(xl, prec)), Xl, prec)
    out = D.ColorSpinorField(mem.empty(F.spinor_bytes(Xl, prec)), Xl, prec)
    ex = comm.HaloExchange(grid, Xl, prec, mode=mode, backend=be, dist=dist)
    for _ in range(reps):  # repeated applications exercise the double-buffered ghost zones
        comm.apply_wilson_distributed(ex, out, din, U, -kappa, dx, parity, dagger)
    mem.sync()
    got = F.spinor_from_native(mem.get(out.buf), F.volume_cb(Xl), prec)
    want = comm.local_slice(ref, Xg, Xl, grid.coords, ("spinor1", parity))
    lvl, dev, _ = oracle.compare_spinor(want, got)
    timed_out = ex.timed_out() if mode == "p2p" else False
    q.put((rank, dev, timed_out))
    dist.barrier()
    dist.destroy_process_group()
