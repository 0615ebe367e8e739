"""Shared multi-rank worker: every rank owns one block of a global lattice, exchanges halos with its neighbours and
applies x + a D in; the result must match the oracle on the GLOBAL lattice.  mode "host": gloo + host twin (CPU tier);
mode "p2p"/"nccl": one GPU per rank, libquda_b200.so, NVLink peer writes / NCCL send-recv (GPU tier)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, grid_dims, Xl, prec, recon, q, mode="host", reps=3, n_src=1):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch
    import torch.distributed as dist
    fused = mode == "p2p-fused"  # the C++ operator layer's schedule: pack + interior + boundary as ONE launch per Dslash
    if fused:
        mode = "p2p"
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
    if n_src > 1:
        # a multi-RHS batch: ONE batched exchange (one pack launch / one message per face for all sources), then every
        # source's Dslash on its own ghost slab; source i is s rolled by i sites so that the sources differ
        ex = comm.HaloExchange(grid, Xl, prec, mode=mode, backend=be, dist=dist, n_src=n_src)
        srcs = [np.roll(s, i, axis=0) for i in range(n_src)]
        dins = [D.ColorSpinorField(mem.put(F.spinor_to_native(comm.local_slice(v, Xg, Xl, grid.coords, ("spinor1", 1 - parity)), prec)), Xl, prec)
                for v in srcs]
        outs = [D.ColorSpinorField(mem.empty(F.spinor_bytes(Xl, prec)), Xl, prec) for _ in range(n_src)]
        for _ in range(reps):
            comm.apply_wilson_distributed(ex, outs, dins, U, -kappa, [dx] * n_src, parity, dagger)
        comm.apply_wilson_distributed(ex, out, din, U, -kappa, dx, parity, dagger)  # a single exchange on the same buffers
        mem.sync()
        dev = 0.0
        for i in range(n_src + 1):
            v, o = (srcs[i], outs[i]) if i < n_src else (s, out)
            r = xs.astype(np.float64) - kappa * oracle.wil_dslash(gauge, v, Xg, parity, dagger).astype(np.float64)
            got = F.spinor_from_native(mem.get(o.buf), F.volume_cb(Xl), prec)
            dev = max(dev, oracle.compare_spinor(comm.local_slice(r, Xg, Xl, grid.coords, ("spinor1", parity)), got)[1])
        q.put((rank, dev, ex.timed_out() if mode == "p2p" else False))
        dist.barrier()
        dist.destroy_process_group()
        return
    ex = comm.HaloExchange(grid, Xl, prec, mode=mode, backend=be, dist=dist)
    if fused:
        from quda_b200 import dirac as DR
        cs = ex.comm_struct()
        op = DR.Dirac("wilson", U, 0.0, comm=cs)
        for _ in range(reps):
            op.DslashXpay(out, din, parity, dx, -kappa, dagger=bool(dagger))
    else:
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


def cg_worker(rank, world, port, grid_dims, Xl, q, mixed=True, kind="wilsonpc"):
    """Distributed CG (normal equations on the even-odd preconditioned operator) through the C++ operator / solver layer:
    every Dslash inside the solver exchanges its halo over NVLink, scalars go through an all-reduce.  The gathered
    solution is verified on the host against the oracle's full operator on the GLOBAL lattice (invert_test's criterion)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import oracle
    from common import CudaMem
    from quda_b200 import comm, dirac as DR, dslash as D, fields as F
    grid = comm.ProcessGrid(grid_dims, rank)
    Xg = [Xl[d] * grid_dims[d] for d in range(4)]
    kappa = 0.12195
    gauge = oracle.random_gauge(Xg, 8, seed=137)
    clover = oracle.random_clover(Xg, 8, seed=138) if "clover" in kind else None
    b = oracle.random_spinor(Xg, 8, seed=77, nparity=2)
    Vhl = F.volume_cb(Xl)

    def local_gauge(prec):
        gl = comm.local_slice(gauge, Xg, Xl, grid.coords, "gauge")
        ghost_from = []
        for d in range(4):
            c = list(grid.coords)
            c[d] = (c[d] - 1) % grid_dims[d]
            ghost_from.append(comm.local_slice(gauge, Xg, Xl, c, "gauge") if grid_dims[d] > 1 else None)
        gbuf, gmeta = F.gauge_to_native(gl, Xl, prec, 12, ghost_from=ghost_from)
        return D.GaugeField(CudaMem.put(gbuf), Xl, prec, 12, gmeta, t_boundary=-1,
                            first_time_slice=grid.first_time_slice(), last_time_slice=grid.last_time_slice())

    def local_clover(prec):
        if clover is None:
            return None
        cl = comm.local_slice(clover, Xg, Xl, grid.coords, "clover")
        cbuf, cmeta = F.clover_to_native(cl, Xl, prec, compressed=True)
        return D.CloverField(CudaMem.put(cbuf), Xl, prec, cmeta, dynamic=True)

    ops, keep = {}, []
    for prec in ((8, 4) if mixed else (8,)):
        ex = comm.HaloExchange(grid, Xl, prec, mode="p2p", dist=dist)
        cs = ex.comm_struct()
        U, A = local_gauge(prec), local_clover(prec)
        ops[prec] = DR.Dirac(kind, U, kappa, clover=A, comm=cs)
        keep += [ex, cs, U, A]
    pc = ops[8]
    # local pieces of b: parity blocks [even | odd]
    bl = np.concatenate([comm.local_slice(b[p * (len(b) // 2):(p + 1) * (len(b) // 2)], Xg, Xl, grid.coords, ("spinor1", p))
                         for p in range(2)])
    pb = F.spinor_bytes(Xl, 8)
    bdev = D.ColorSpinorField(CudaMem.put(np.concatenate([F.spinor_to_native(bl[p * Vhl:(p + 1) * Vhl], 8) for p in range(2)])), Xl, 8, 2)
    xdev = D.ColorSpinorField(CudaMem.empty(2 * pb), Xl, 8, 2)
    src_p, sol_p = pc.prepare(xdev, bdev)
    src = D.ColorSpinorField(xdev.buf[src_p * pb:(src_p + 1) * pb], Xl, 8)
    sol = D.ColorSpinorField(xdev.buf[sol_p * pb:(sol_p + 1) * pb], Xl, 8)
    rhs = D.ColorSpinorField(CudaMem.empty(pb), Xl, 8)
    pc.Mdag(rhs, src)
    sol.buf.zero_()
    res = DR.invert_cg(pc, ops.get(4), sol, rhs, tol=1e-10, maxiter=3000)
    pc.reconstruct(xdev, bdev)
    torch.cuda.synchronize()
    raw = CudaMem.get(xdev.buf)
    xl = np.concatenate([F.spinor_from_native(raw[p * pb:(p + 1) * pb], Vhl, 8) for p in range(2)])
    # gather the solution: every rank contributes its block to the global field
    xg = np.zeros_like(b, dtype=np.float64)
    Vhg = F.volume_cb(Xg)
    blocks = [None] * world
    dist.all_gather_object(blocks, (grid.coords, xl))
    for coords, blk in blocks:
        off = np.array([coords[d] * Xl[d] for d in range(4)])
        for p in range(2):
            idx = F.cb_index(F.cb_coords(Xl, p) + off, Xg)
            xg[p * Vhg + idx] = blk[p * Vhl:(p + 1) * Vhl]
    if clover is None:
        Mx = oracle.wil_mat(gauge, xg, Xg, kappa, 0)
    else:
        Mx = oracle.clover_mat(gauge, clover, xg, Xg, kappa, 0)
    true_res = float(np.linalg.norm(Mx.ravel() - b.ravel()) / np.linalg.norm(b.ravel()))
    q.put((rank, res.iter, res.reliable_updates, res.true_res, true_res, any(e.timed_out() for e in keep if hasattr(e, "timed_out"))))
    dist.barrier()
    dist.destroy_process_group()
