"""CPU tier, world_size = 2 over gloo: process-grid topology, local slicing of global fields, face packing,
neighbour exchange and the interior + exterior split -- each rank computes its half of a global lattice with
the host twin and the gathered result must match the oracle on the GLOBAL lattice."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, grid_dims, Xl, prec, recon, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from common import HostMem, assert_close, twin_backend
    from quda_b200 import comm, dslash as D, fields as F
    be = twin_backend()
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
    U = D.GaugeField(gbuf, Xl, prec, recon, gmeta, t_boundary=-1, first_time_slice=grid.first_time_slice(),
                     last_time_slice=grid.last_time_slice())
    sl = comm.local_slice(s, Xg, Xl, grid.coords, ("spinor1", 1 - parity))
    xl = comm.local_slice(xs, Xg, Xl, grid.coords, ("spinor1", parity))
    din = D.ColorSpinorField(F.spinor_to_native(sl, prec), Xl, prec)
    dx = D.ColorSpinorField(F.spinor_to_native(xl, prec), Xl, prec)
    out = D.ColorSpinorField(np.zeros(F.spinor_bytes(Xl, prec), dtype=np.uint8), Xl, prec)
    ex = comm.HaloExchange(grid, Xl, prec, mode="host", backend=be, dist=dist)
    for _ in range(3):  # repeated applications exercise the double-buffered ghost zones
        comm.apply_wilson_distributed(ex, out, din, U, -kappa, dx, parity, dagger)
    got = F.spinor_from_native(out.buf, F.volume_cb(Xl), prec)
    want = comm.local_slice(ref, Xg, Xl, grid.coords, ("spinor1", parity))
    lvl, dev, _ = oracle.compare_spinor(want, got)
    q.put((rank, dev))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("grid_dims,Xl", [((1, 1, 1, 2), (4, 4, 4, 4)), ((2, 1, 1, 1), (4, 4, 6, 4)), ((1, 2, 1, 1), (4, 4, 4, 6))])
@pytest.mark.parametrize("prec,recon", [(8, 18), (4, 12), (2, 8)])
def test_two_rank_wilson_matches_global_oracle(grid_dims, Xl, prec, recon):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, grid_dims, Xl, prec, recon, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tol = {8: 1e-11, 4: 1e-4, 2: 1e-2}[prec]
    for rank, dev in res:
        assert dev <= tol, (rank, dev)


def test_process_grid_topology():
    from quda_b200.comm import ProcessGrid
    g = ProcessGrid((1, 2, 2, 2), 5)
    assert g.coords == [0, 1, 0, 1] and g.rank_of(g.coords) == 5
    assert g.neighbor(3, +1) == g.rank_of([0, 1, 0, 0]) and g.neighbor(1, -1) == g.rank_of([0, 0, 0, 1])
    assert ProcessGrid.default_dims(8) == [1, 2, 2, 2] and ProcessGrid.default_dims(2) == [1, 1, 1, 2]
    assert g.comm_dim() == [0, 1, 1, 1]
