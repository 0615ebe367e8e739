"""GPU tier at BASELINE.json's full single-GPU size (32^4): the configurations the metric is quoted on, compared directly
with the CPU oracle on the same seeded fields (the oracle needs ~1 s per 32^4 Dslash), plus the size-independent
properties of the operator -- linearity and the adjoint identity <w, D v> = <D^dagger w, v> -- which catch indexing
errors that only show up at production tile counts (8192 CTAs, multi-wave launches, the bench.py launch shape)."""
import numpy as np
import pytest

import oracle
from common import CudaMem, assert_close
from quda_b200 import dslash as D
from quda_b200 import fields as F

pytestmark = pytest.mark.gpu
X = (32, 32, 32, 32)
VH = 32 ** 4 // 2


class Fields:
    """host fields in oracle order + device-marshaled native images (b200_copy_gauge / b200_copy_clover)"""

    def __init__(self, host_prec):
        self.hp = host_prec
        self.gauge = oracle.random_gauge(X, host_prec, seed=137)  # anti-periodic in t, as dslash_test does
        self.s = oracle.random_spinor(X, host_prec, seed=5)
        self.w = oracle.random_spinor(X, host_prec, seed=6)

    def U(self, prec, recon):
        return D.load_gauge(self.gauge, X, prec, recon, t_boundary=-1)

    @staticmethod
    def dev(host, prec):
        return D.ColorSpinorField(CudaMem.put(F.spinor_to_native(host, prec)), X, prec, 1)

    @staticmethod
    def empty(prec):
        return D.ColorSpinorField(CudaMem.empty(F.spinor_bytes(X, prec)), X, prec, 1)

    @staticmethod
    def host(field):
        CudaMem.sync()
        return F.spinor_from_native(CudaMem.get(field.buf), VH, field.prec)


@pytest.fixture(scope="module")
def f32():
    return Fields(4)


def test_config2_fp32_recon12_vs_oracle_and_properties(f32):
    """BASELINE config 2: 32^4, fp32, recon-12, single parity, no xpay -- the bench.py workload"""
    U = f32.U(4, 12)
    s_dev, out = f32.dev(f32.s, 4), f32.empty(4)
    D.ApplyWilson(out, s_dev, U, 0.0, None, 0, 0)
    Ds = f32.host(out)
    assert_close(oracle.wil_dslash(f32.gauge, f32.s, X, 0, 0), Ds, 4, 12, "32^4 fp32 recon-12 vs oracle")
    # linearity: D (2 s - 0.5 w) = 2 D s - 0.5 D w
    D.ApplyWilson(out, f32.dev(f32.w, 4), U, 0.0, None, 0, 0)
    Dw = f32.host(out)
    D.ApplyWilson(out, f32.dev((2.0 * f32.s - 0.5 * f32.w).astype(np.float32), 4), U, 0.0, None, 0, 0)
    lin = f32.host(out)
    scale = np.abs(Ds).max()
    assert np.abs(lin - (2.0 * Ds - 0.5 * Dw)).max() < 2e-5 * scale
    # adjoint identity with the dagger flag (w lives on the output parity): <w, D s> = <D^dagger w, s>
    D.ApplyWilson(out, f32.dev(f32.w, 4), U, 0.0, None, 1, 1)
    Ddag_w = f32.host(out).astype(np.float64)
    lhs = np.vdot(f32.w.astype(np.float64).ravel(), Ds.astype(np.float64).ravel())
    rhs = np.vdot(Ddag_w.ravel(), f32.s.astype(np.float64).ravel())
    assert abs(lhs - rhs) < 1e-5 * abs(lhs)


def test_fp64_recon18_32cubed_vs_oracle():
    """fp64 at full size against the fp64 oracle: the north star's 1e-12 and the reference's 1e-11 gate"""
    fd = Fields(8)
    U = fd.U(8, 18)
    out = fd.empty(8)
    D.ApplyWilson(out, fd.dev(fd.s, 8), U, 0.0, None, 1, 1)
    got = fd.host(out)
    ref = oracle.wil_dslash(fd.gauge, fd.s, X, 1, 1)
    assert_close(ref, got, 8, 18, "32^4 fp64 recon-18 vs oracle")
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_config3_clover_pc_half_recon8_vs_oracle(f32):
    """BASELINE config 3: clover-preconditioned Dslash, half precision, recon-8, 32^4 (tolerance 1e-3 x 10,
    tests/dslash_test.cpp:74-76)"""
    clover = oracle.random_clover(X, 4, seed=11)
    clover_inv = oracle.clover_invert(clover)
    U = f32.U(2, 8)
    A = D.load_clover(clover, X, 2, compressed=True, dynamic=True)
    out = f32.empty(2)
    D.ApplyWilsonCloverPreconditioned(out, f32.dev(f32.s, 2), U, A, 0.0, None, 0, 0)
    ref = oracle.clover_dslash(f32.gauge, clover_inv, f32.s, X, 0, 0)
    assert_close(ref, f32.host(out), 2, 8, "32^4 clover-pc half recon-8 vs oracle")
