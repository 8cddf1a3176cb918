"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/titgpu.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import titsolver_b200 as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "titgpu.h")).read()
    return sorted(set(re.findall(r"TITGPU_API[^;]*?\b(titgpu_\w+)\s*\(", src)))


def test_header_symbols_exported():
    lib = tb.load_library()
    syms = declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), s
    assert set(tb.ABI_SYMBOLS) == set(syms)


def test_version_and_error_path_without_compute():
    lib = tb.load_library()
    assert b"titgpu" in lib.titgpu_version()
    # A bad (dim, kernel) is rejected before any CUDA call is made.
    h = ctypes.c_void_p()
    rc = lib.titgpu_create(ctypes.byref(h), 0, 7, 4, 0, 3)
    assert rc != 0
    assert b"no engine" in lib.titgpu_last_error(h)
    lib.titgpu_destroy(h)


def test_product_does_not_reference_oracle():
    """The product path must never route through oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "titsolver_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "liboracle" not in txt, f
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("the oracle", ""), f
