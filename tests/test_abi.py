"""CPU checks of the C-ABI boundary: the library loads and exports every symbol include/rdb200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest


def _declared_symbols():
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "rdb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rdb_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported(rd):
    lib = ctypes.CDLL(rd.engine.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"librdb200.so does not export {n}"
    assert sorted(rd.engine.EXPORTED_SYMBOLS) == names      # the ctypes binding covers the whole header


def test_no_cpu_fallback(rd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rd.engine.EngineError, match="no CPU fallback"):
        rd.engine.Context(0)


def test_oracle_is_not_imported_by_the_product(rd):
    root = os.path.join(os.path.dirname(__file__), "..", "reversediff.jl_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
