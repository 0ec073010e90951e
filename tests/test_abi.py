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
