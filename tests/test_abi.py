"""The C-ABI library loads on a box without a GPU and exports every symbol that
include/mpvss_b200.h declares; without a device the context cannot be created (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "mpvss_b200.h")).read()
    return sorted(set(re.findall(r"\b(mpvss_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    from mpvss_rs_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libmpvss_b200.so not built (run make / __graft_entry__.build())")
    so = ctypes.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(so, n), f"{n} declared in include/mpvss_b200.h but not exported"
    assert set(lib.SIGNATURES) == set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mpvss_rs_b200 as m
    with pytest.raises(m.MpvssError):
        m.Group("modp")


def test_header_compiles_as_plain_c_and_links():
    """include/mpvss_b200.h is a C header (no C++), and a plain-C program links against the library --
    the shape every foreign binding (Rust FFI, cgo, JNI) depends on.  Running it needs a GPU (next test)."""
    import subprocess
    from mpvss_rs_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        pytest.skip("libmpvss_b200.so not built")
    exe = os.path.join(ROOT, "tests", "c_abi", "smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "smoke.c"), "-L", os.path.dirname(lib.LIB_PATH),
                           "-lmpvss_b200", "-Wl,-rpath," + os.path.dirname(lib.LIB_PATH), "-o", exe])
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_plain_c_program_runs_the_hot_path():
    import subprocess
    test_header_compiles_as_plain_c_and_links()
    out = subprocess.run([os.path.join(ROOT, "tests", "c_abi", "smoke")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "c abi smoke ok" in out.stdout, out.stdout + out.stderr
