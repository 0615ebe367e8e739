"""GPU tier (-m gpu): the TMA-staged marching kernel (quda_b200/csrc/tma_kernel.cuh: cp.async.bulk.tensor box loads,
mbarrier producer/consumer pipeline) through the C ABI against the CPU oracle.  B200_TMA=2 makes the library fail
instead of silently falling back to the gather kernel, so every case below is a real TMA launch."""
import numpy as np
import pytest

import ops
import oracle
from common import CudaMem, Problem, assert_close
from quda_b200 import dslash as D
from quda_b200 import lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture
def tma(monkeypatch):
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    monkeypatch.setenv("B200_TMA", "2")
    lib = L.load()
    before = lib.b200_launch_count()
    yield monkeypatch
    assert lib.b200_launch_count() > before, "no kernel from libquda_b200.so was launched"


@pytest.mark.parametrize("X,prec,recon", [((16, 4, 4, 4), 4, 12), ((16, 8, 4, 6), 4, 12), ((8, 4, 8, 4), 4, 8),
                                          ((32, 4, 4, 4), 4, 18), ((16, 4, 8, 4), 8, 12), ((16, 4, 4, 6), 8, 18),
                                          ((8, 6, 10, 4), 8, 8), ((48, 2, 2, 4), 4, 12), ((16, 16, 16, 16), 4, 12),
                                          ((16, 16, 16, 16), 8, 12)])
def test_tma_wilson_dslash(tma, X, prec, recon):
    P = Problem(X, prec, recon, CudaMem)
    for parity in (0, 1):
        for dagger in (0, 1):
            s = P.spinor(seed=11 + parity)
            out = P.empty()
            D.ApplyWilson(out, P.to_dev(s), P.U, 0.0, None, parity, dagger)
            ref = oracle.wil_dslash(P.gauge, s, X, parity, dagger)
            dev = assert_close(ref, P.to_host(out), prec, recon, f"X={X} parity={parity} dagger={dagger}")
            if prec == 8:
                assert dev <= 1e-12


@pytest.mark.parametrize("links", [-1, 0, 2, 3, 4])
@pytest.mark.parametrize("grid", [1, 3, 37, 148, 1000])
def test_tma_pipeline_depths_and_work_ranges(tma, links, grid):
    """both link modes (-1: register stream, 0 / n: shared-memory stages) and every work partition (ranges that start / end
    mid-tile, one item per CTA, one CTA for everything): no deadlock, no slot hazard, every site visited once"""
    tma.setenv("B200_TMA_LINKS", str(links))
    tma.setenv("B200_TMA_GRID", str(grid))
    X = (16, 4, 8, 6)
    P = Problem(X, 4, 12, CudaMem, anisotropy=1.7)
    s = P.spinor(seed=5)
    out = P.empty()
    D.ApplyWilson(out, P.to_dev(s), P.U, 0.0, None, 1, 0)
    assert_close(oracle.wil_dslash(P.gauge, s, X, 1, 0), P.to_host(out), 4, 12, f"links={links} grid={grid}")


@pytest.mark.parametrize("prefetch", [2, 3, 4])
@pytest.mark.parametrize("rings", ["4 2", "5 2", "6 3", "8 4"])
@pytest.mark.parametrize("grid", [5, 148])
def test_tma_register_stream_links(tma, prefetch, rings, grid):
    """register-stream links: every prefetch distance (the reload of a direction pair targets this item or the next one)
    and spinor ring depth, with work ranges that cross tiles and parities (full field)"""
    tma.setenv("B200_TMA_LINKS", "-1")
    tma.setenv("B200_TMA_PREFETCH", str(prefetch))
    tma.setenv("B200_TMA_RINGS", rings)
    tma.setenv("B200_TMA_GRID", str(grid))
    X = (8, 4, 4, 6)
    P = Problem(X, 4, 12, CudaMem)
    full = P.spinor(seed=3, nparity=2)
    out = P.empty(2)
    D.ApplyWilson(out, P.to_dev(full, 2), P.U, 0.0, None, D.QUDA_INVALID_PARITY, 1)
    ref = np.concatenate([oracle.wil_dslash(P.gauge, full[P.Vh:], X, 0, 1), oracle.wil_dslash(P.gauge, full[:P.Vh], X, 1, 1)])
    assert_close(ref, P.to_host(out), 4, 12, f"prefetch={prefetch} rings={rings}")


@pytest.mark.parametrize("l2pf", [-1, 1, 2, 5])
def test_tma_l2_prefetch_lookahead(tma, l2pf):
    """the L2 prefetch of the link boxes is a hint (no functional effect) but its coordinates come from the same box
    function: every look-ahead, including one beyond the end of a CTA's work range, must leave the result unchanged"""
    tma.setenv("B200_TMA_L2PF", str(l2pf))
    tma.setenv("B200_TMA_GRID", "7")
    X = (16, 4, 4, 6)
    P = Problem(X, 4, 12, CudaMem)
    s = P.spinor(seed=8)
    out = P.empty()
    D.ApplyWilson(out, P.to_dev(s), P.U, 0.0, None, 0, 0)
    assert_close(oracle.wil_dslash(P.gauge, s, X, 0, 0), P.to_host(out), 4, 12, f"l2pf={l2pf}")


@pytest.mark.parametrize("tile", ["1 1", "2 1", "1 2", "4 2", "2 4", "4 4", "8 1"])
def test_tma_tiles(tma, tile):
    tma.setenv("B200_TMA_TILE", tile)
    X = (8, 8, 4, 4)
    P = Problem(X, 4, 12, CudaMem)
    s = P.spinor(seed=6)
    for parity in (0, 1):
        out = P.empty()
        D.ApplyWilson(out, P.to_dev(s), P.U, 0.0, None, parity, 1)
        assert_close(oracle.wil_dslash(P.gauge, s, X, parity, 1), P.to_host(out), 4, 12, f"tile {tile} parity {parity}")


@pytest.mark.parametrize("prec,recon", [(8, 18), (8, 12), (4, 12), (4, 8)])
def test_tma_xpay_fullfield(tma, prec, recon):
    ops.check_xpay_fullfield(CudaMem, None, prec, recon, X=(8, 4, 4, 6))


@pytest.mark.parametrize("prec", [8, 4])
@pytest.mark.parametrize("compressed,dynamic", [(True, True), (False, False)])
def test_tma_clover(monkeypatch, prec, compressed, dynamic):
    monkeypatch.setenv("B200_TMA", "1")
    ops.check_clover(CudaMem, None, prec, 12, compressed, dynamic, X=(8, 4, 6, 4))


def test_tma_matches_gather_kernel_bitwise(monkeypatch):
    """same arithmetic, operation for operation: the TMA kernel and the gather kernel must agree bit for bit"""
    import torch
    X = (16, 8, 8, 8)
    P = Problem(X, 4, 12, CudaMem)
    s = P.to_dev(P.spinor(seed=9))
    outs = []
    for mode in ("0", "2"):
        monkeypatch.setenv("B200_TMA", mode)
        out = P.empty()
        D.ApplyWilson(out, s, P.U, 0.0, None, 0, 0)
        torch.cuda.synchronize()
        outs.append(out.buf.clone())
    assert torch.equal(outs[0], outs[1])


def test_tma_32cubed_fp32_recon12_vs_oracle(tma):
    """BASELINE config 2 at full size through the TMA kernel"""
    from quda_b200 import fields as F
    X = (32, 32, 32, 32)
    gauge = oracle.random_gauge(X, 4, seed=137)
    s = oracle.random_spinor(X, 4, seed=5)
    U = D.load_gauge(gauge, X, 4, 12, t_boundary=-1)
    out = D.ColorSpinorField(CudaMem.empty(F.spinor_bytes(X, 4)), X, 4, 1)
    D.ApplyWilson(out, D.ColorSpinorField(CudaMem.put(F.spinor_to_native(s, 4)), X, 4, 1), U, 0.0, None, 0, 0)
    CudaMem.sync()
    got = F.spinor_from_native(CudaMem.get(out.buf), 32 ** 4 // 2, 4)
    assert_close(oracle.wil_dslash(gauge, s, X, 0, 0), got, 4, 12, "32^4 fp32 recon-12 TMA vs oracle")
