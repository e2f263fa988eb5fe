"""CPU: the C-ABI library builds in-tree, loads, and exports exactly what include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(obia_b200_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from obia_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signatures and the header disagree"
    assert lib.obia_b200_version() >= 100
    assert isinstance(lib.obia_b200_launch_count(), int)


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device (no compute on CPU boxes)."""
    from obia_b200 import _lib
    lib = _lib.load()
    assert lib.obia_b200_slic_workspace_bytes(0, 10, 3, 4, 2, 2) == -1
    assert lib.obia_b200_slic_workspace_bytes(100, 100, 3, 16, 25, 25) > 0
    assert lib.obia_b200_connectivity_workspace_bytes(100, 100) >= 100 * 100 * 29
    assert lib.obia_b200_zonal_workspace_bytes(10, 3) > 0
    rc = lib.obia_b200_enforce_connectivity(None, None, None, 10, 10, 1, 5, 1, None, None)
    assert rc == -1 and b"bad argument" in lib.obia_b200_last_error()
    rc = lib.obia_b200_band_minmax(None, 0, 3, None, None, None, None)
    assert rc == -1
    # pointers that would make the 128-bit loads fault are rejected, not dereferenced
    rc = lib.obia_b200_band_minmax(ctypes.c_void_p(0x1008), 10, 3, None, ctypes.c_void_p(0x2000),
                                   ctypes.c_void_p(0x3000), None)
    assert rc == -1 and b"aligned" in lib.obia_b200_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from obia_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libobia_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may touch oracle/."""
    pat = re.compile(r"^\s*(import|from)\s+(slic_oracle|stats_oracle|tiling_oracle|oracle)\b", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "obia_b200")):
        for fn in files:
            if fn.endswith(".py"):
                text = open(os.path.join(dirpath, fn)).read()
                assert not pat.search(text), f"{fn} imports the oracle"
