"""CPU tier: the host twin rebuilt with AddressSanitizer + UndefinedBehaviorSanitizer runs a cross-section of the twin tests
in a subprocess (LD_PRELOAD=libasan).  The twin executes the product's own site code, index maps, ghost / pad addressing and
launch-geometry code (quda_b200/csrc/*.h), so an out-of-bounds access or signed overflow found here is one the CUDA kernels
would commit on the GPU -- the CPU stand-in for compute-sanitizer, which needs a device."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TWIN = os.path.join(HERE, "hosttwin")


def _runtime(name):
    p = subprocess.run(["gcc", "-print-file-name=" + name], capture_output=True, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


def test_twin_suite_is_clean_under_asan_and_ubsan():
    asan = _runtime("libasan.so")
    if asan is None:
        pytest.skip("no libasan in this toolchain")
    # the sanitizer build of this template-heavy translation unit takes ~15 minutes: run where it is already built and up to
    # date (`make -C tests/hosttwin asan`), or on request (B200_SANITIZE=1)
    if os.environ.get("B200_SANITIZE") == "1":
        subprocess.check_call(["make", "-s", "-C", TWIN, "asan"])
    elif subprocess.run(["make", "-q", "-C", TWIN, "asan"]).returncode != 0:
        pytest.skip("tests/hosttwin/_build/libhosttwin_asan.so not built / stale (make -C tests/hosttwin asan, ~15 min; or B200_SANITIZE=1)")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1",
               B200_TWIN_LIB=os.path.join(TWIN, "_build", "libhosttwin_asan.so"), OMP_NUM_THREADS="4")
    sel = "all_masks or precisions or batched or fused or site_granular or minimal or multi_rhs_wilson or clover or twisted"
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-s", "-p", "no:cacheprovider", os.path.join(HERE, "test_twin_wilson.py"),
                        os.path.join(HERE, "test_twin_ops.py"), os.path.join(HERE, "test_twin_tma.py"), "-k", sel],
                       capture_output=True, text=True, env=env, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert "AddressSanitizer" not in r.stdout + r.stderr and "runtime error" not in r.stdout + r.stderr, tail
    assert r.returncode == 0 and " passed" in r.stdout, tail
