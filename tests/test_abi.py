"""The C-ABI library loads and exports every symbol include/dbsp_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import dbsp_b200
from dbsp_b200._capi import CApi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dbsp_b200.h")).read()
    return sorted(set(re.findall(r"\b(dbsp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as ge

    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    # and the Python binding covers every one of them
    assert sorted(CApi.symbols("dbsp_")) == syms


def test_no_cpu_fallback():
    """Without a device the product backend must fail loudly, not fall back."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dbsp_b200.runtime import Runtime
    from dbsp_b200._capi import DbspError

    with pytest.raises(DbspError, match="NO_DEVICE"):
        Runtime(0)
