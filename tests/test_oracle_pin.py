"""Pin the CPU oracle (oracle/wilson_oracle.c) against the reference's OWN host sources compiled in place
(oracle/_ref/libquda_hostref.so: tests/host_reference/{wilson_dslash,clover}_reference.cpp,
tests/utils/{host_utils,index_utils,host_blas}.cpp).  The reference tree stores no golden vectors
 so this is what anchors parity.  Same compiler flags on both sides -> bit-identical."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")

DIMS = [(4, 4, 4, 4), (8, 4, 6, 2), (2, 4, 6, 8), (8, 8, 8, 8)]


@pytest.mark.parametrize("X", DIMS)
@pytest.mark.parametrize("prec", [8, 4])
def test_field_generators_bit_identical(X, prec):
    R = oracle.Reference(X)
    for aniso, apbc in [(1.0, True), (2.38, False)]:
        g_ref = R.random_gauge(prec, seed=137, anisotropy=aniso, antiperiodic_t=apbc)
        g_orc = oracle.random_gauge(X, prec, seed=137, anisotropy=aniso, antiperiodic_t=apbc)
        assert np.array_equal(g_ref, g_orc)
    c_ref = R.random_clover(prec, seed=5)
    c_orc = oracle.random_clover(X, prec, seed=5)
    assert np.array_equal(c_ref, c_orc)


def test_random_gauge_is_su3():
    X = (4, 4, 4, 4)
    g = oracle.random_gauge(X, 8, antiperiodic_t=False)
    u = g[..., 0] + 1j * g[..., 1]
    uu = np.einsum("dvab,dvcb->dvac", u, u.conj())
    assert np.allclose(uu, np.eye(3), atol=1e-12)
    assert np.allclose(np.linalg.det(u), 1.0, atol=1e-12)


@pytest.mark.parametrize("X", DIMS)
@pytest.mark.parametrize("prec", [8, 4])
def test_wilson_operators_bit_identical(X, prec):
    R = oracle.Reference(X)
    g = oracle.random_gauge(X, prec, seed=137)
    s = oracle.random_spinor(X, prec, seed=137)
    for parity in (0, 1):
        for dagger in (0, 1):
            assert np.array_equal(R.wil_dslash(g, s, parity, dagger), oracle.wil_dslash(g, s, X, parity, dagger))
    kappa = 0.12195
    for matpc in (oracle.MATPC_EVEN_EVEN, oracle.MATPC_ODD_ODD):
        for dagger in (0, 1):
            assert np.array_equal(R.wil_matpc(g, s, kappa, matpc, dagger), oracle.wil_matpc(g, s, X, kappa, matpc, dagger))
    full = oracle.random_spinor(X, prec, seed=99, nparity=2)
    for dagger in (0, 1):
        assert np.array_equal(R.wil_mat(g, full, kappa, dagger), oracle.wil_mat(g, full, X, kappa, dagger))


@pytest.mark.parametrize("X", DIMS[:3])
@pytest.mark.parametrize("prec", [8, 4])
def test_clover_operators_bit_identical(X, prec):
    R = oracle.Reference(X)
    g = oracle.random_gauge(X, prec, seed=137)
    c = oracle.random_clover(X, prec, seed=11)
    cinv = oracle.clover_invert(c)
    s = oracle.random_spinor(X, prec, seed=137)
    kappa = 0.12195
    for parity in (0, 1):
        assert np.array_equal(R.apply_clover(c, s, parity), oracle.apply_clover(c, s, X, parity))
        for dagger in (0, 1):
            assert np.array_equal(R.clover_dslash(g, cinv, s, parity, dagger), oracle.clover_dslash(g, cinv, s, X, parity, dagger))
    for matpc in range(4):
        for dagger in (0, 1):
            a = R.clover_matpc(g, c, cinv, s, kappa, matpc, dagger)
            b = oracle.clover_matpc(g, c, cinv, s, X, kappa, matpc, dagger)
            assert np.array_equal(a, b), (matpc, dagger)
    full = oracle.random_spinor(X, prec, seed=99, nparity=2)
    for dagger in (0, 1):
        assert np.array_equal(R.clover_mat(g, c, full, kappa, dagger), oracle.clover_mat(g, c, full, X, kappa, dagger))


@pytest.mark.parametrize("prec", [8, 4])
def test_clover_inverse_is_inverse(prec):
    """The reference inverts the clover term on the device; our host inverse is checked through the
    reference's own apply_clover: A^{-1} (A v) == v."""
    X = (4, 4, 4, 4)
    R = oracle.Reference(X)
    c = oracle.random_clover(X, prec, seed=3)
    cinv = oracle.clover_invert(c)
    s = oracle.random_spinor(X, prec, seed=1)
    for parity in (0, 1):
        back = R.apply_clover(cinv, R.apply_clover(c, s, parity), parity)
        assert np.allclose(back, s, rtol=0, atol=1e-13 if prec == 8 else 2e-6)


def test_compare_spinor_metric():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(24 * 512)
    lvl, dev, _ = oracle.compare_spinor(a, a)
    assert lvl == 16
    b = a.copy()
    b[7] += 3e-6 * np.abs(a).max()
    lvl, dev, fails = oracle.compare_spinor(a, b)
    assert lvl == 5 and fails[5] == 1 and fails[4] == 0


@pytest.mark.parametrize("X", DIMS[:3])
@pytest.mark.parametrize("prec", [8, 4])
def test_twisted_mass_operators_bit_identical(X, prec):
    """twist_gamma5 / tm_dslash / tm_mat / tm_matpc (wilson_dslash_reference.cpp:139-315), singlet flavour, all four
    matpc types, both parities, dagger"""
    R = oracle.Reference(X)
    g = oracle.random_gauge(X, prec, seed=137)
    s = oracle.random_spinor(X, prec, seed=21)
    full = oracle.random_spinor(X, prec, seed=22, nparity=2)
    kappa, mu = 0.12195, 0.1
    for dagger in (0, 1):
        for inverse in (False, True):
            assert np.array_equal(R.twist_gamma5(s, kappa, mu, dagger, inverse), oracle.twist_gamma5(s, kappa, mu, dagger, inverse))
        for matpc in range(4):
            for parity in (0, 1):
                assert np.array_equal(R.tm_dslash(g, s, kappa, mu, parity, dagger, matpc),
                                      oracle.tm_dslash(g, s, X, kappa, mu, parity, dagger, matpc))
            assert np.array_equal(R.tm_matpc(g, s, kappa, mu, matpc, dagger), oracle.tm_matpc(g, s, X, kappa, mu, matpc, dagger))
        assert np.array_equal(R.tm_mat(g, full, kappa, mu, dagger), oracle.tm_mat(g, full, X, kappa, mu, dagger))
