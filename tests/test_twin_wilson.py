"""CPU tier: the product's site code (compiled for the host, tests/hosttwin) against the oracle.
Same b200_dslash_args ABI, same native layouts, same argument validation as the CUDA library."""
import numpy as np
import pytest

import oracle
from common import HostMem, Problem, assert_close, twin_backend
from quda_b200 import dslash as D

PRECS = [8, 4, 2]
RECONS = [18, 12, 8]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("recon", RECONS)
@pytest.mark.parametrize("X", [(4, 4, 4, 4), (8, 4, 6, 2)])
def test_wilson_dslash(prec, recon, X):
    be = twin_backend()
    P = Problem(X, prec, recon, HostMem)
    for parity in (0, 1):
        for dagger in (0, 1):
            s = P.spinor(seed=11 + parity)
            out = P.empty()
            D.ApplyWilson(out, P.to_dev(s), P.U, 0.0, None, parity, dagger, backend=be)
            ref = oracle.wil_dslash(P.gauge, s, X, parity, dagger)
            assert_close(ref, P.to_host(out), prec, recon, f"dslash parity={parity} dagger={dagger}")
