"""CPU tier: the C-ABI library loads, exports every symbol include/b200_dslash.h declares, and refuses to
compute without a GPU (no CPU fallback in the product)."""
import ctypes as C
import os
import re

import pytest

from quda_b200 import dslash as D
from quda_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200_dslash.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = declared_symbols()
    assert set(syms) == set(L.EXPORTED_SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.b200_abi_version() == L.ABI_VERSION


def test_struct_sizes_match_header():
    """ctypes mirrors vs. the C compiler's view of include/b200_dslash.h"""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "b200_dslash.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(b200_spinor),sizeof(b200_gauge),sizeof(b200_clover),sizeof(b200_halo),sizeof(b200_dslash_args),' \
          'sizeof(b200_pack_args), offsetof(b200_dslash_args, halo), sizeof(b200_comm), offsetof(b200_comm, reduce_peer),' \
          'offsetof(b200_comm, reduce_seq), sizeof(b200_solver_param));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write("#include <stddef.h>\n" + src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "t.c"), "-o", os.path.join(td, "t")])
        got = [int(v) for v in subprocess.check_output([os.path.join(td, "t")]).split()]
    want = [C.sizeof(L.Spinor), C.sizeof(L.Gauge), C.sizeof(L.Clover), C.sizeof(L.Halo), C.sizeof(L.DslashArgs),
            C.sizeof(L.PackArgs), L.DslashArgs.halo.offset, C.sizeof(L.Comm), L.Comm.reduce_peer.offset,
            L.Comm.reduce_seq.offset, C.sizeof(L.SolverParam)]
    assert L.REDUCE_MAILBOX_BYTES == 2 * L.MAX_RANKS * 64
    assert got == want


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    args = L.DslashArgs()
    args.abi_version = L.ABI_VERSION
    rc = lib.b200_dslash_apply(C.byref(args))
    assert rc == -4  # B200_ERR_NO_DEVICE
    assert b"no CPU path" in lib.b200_last_error()
    # every compute entry point refuses the same way (none of them can reach a host code path)
    import numpy as np
    buf = np.zeros(64, dtype=np.uint8)
    sp = L.Spinor(buf.ctypes.data, None, 0, 1, 1)
    sps = (L.Spinor * 2)(sp, sp)
    cl = L.Clover(buf.ctypes.data, 0, 1, 1, 1.0, 1.0)
    assert lib.b200_dslash_apply_multi(C.byref(args), 2, sps, sps, None) == -4
    assert lib.b200_clover_apply(C.byref(sp), C.byref(sp), C.byref(cl), 4, 0, 0, None) == -4
    assert lib.b200_twist_gamma5(C.byref(sp), C.byref(sp), 4, 0.1, 0.1, 0, 0, None) == -4
    assert lib.b200_copy_spinor(C.byref(sp), 4, buf.ctypes.data, 4, 1, None) == -4
    pa = L.PackArgs()
    pa.abi_version = L.ABI_VERSION
    pa.in_ = sp
    assert lib.b200_pack_ghost(C.byref(pa)) == -4


def test_ghost_face_bytes():
    lib = L.load()
    X = (C.c_int * 4)(32, 32, 32, 32)
    assert lib.b200_ghost_face_bytes(4, X, 0) == 2 * 786432      # 32^4 fp32 x-face, per parity 786 432 B
    assert lib.b200_ghost_face_bytes(8, X, 3) == 2 * 1572864
    assert lib.b200_ghost_face_bytes(2, X, 1) == 2 * 458752
