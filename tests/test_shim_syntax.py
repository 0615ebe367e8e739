"""CPU tier: the reference-side binding shim/quda_apply_shim.cpp (the translation unit a QUDA maintainer compiles instead of
lib/dslash_wilson*.cu, lib/dslash_twisted_mass*.cu, lib/dslash_clover_helper.cu and lib/dslash_pack2.cu) must parse and
type-check against the REFERENCE's own headers: every QUDA type, accessor and signature it uses is the real one.
Recipe as oracle/Makefile (the two generated headers come from `make -C oracle ref`).  Skipped where /root/reference is
not mounted (the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include")), reason="reference tree not mounted")
def test_shim_type_checks_against_reference_headers():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/gen/quda_define.h"])
    inc = [os.path.join(ROOT, "oracle", "_ref", "gen"), f"{REF}/include", f"{REF}/include/targets/cuda",
           f"{REF}/include/targets/generic", f"{REF}/include/targets/cuda/externals", f"{REF}/include/externals", f"{REF}/lib",
           "/usr/local/cuda/include", os.path.join(ROOT, "include")]
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w"] + [f"-I{i}" for i in inc] + [os.path.join(ROOT, "shim", "quda_apply_shim.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_shim_defines_every_replaced_entry_point():
    src = open(os.path.join(ROOT, "shim", "quda_apply_shim.cpp")).read()
    for sym in ("void ApplyWilson(", "void ApplyWilsonClover(", "void ApplyWilsonCloverPreconditioned(", "void ApplyTwistedMass(",
                "void ApplyTwistedMassPreconditioned(", "void ApplyClover(", "void PackGhost(", "static void fill_halo(",
                "static void apply_partitioned("):
        assert sym in src, sym
