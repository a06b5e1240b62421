"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares,
and compute entry points fail loudly (no CPU fallback) when no GPU is present."""
import ctypes
import glob
import os
import re
import pytest
import plslam_b200 as pl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = []
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
        syms += re.findall(r"\b(pl_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(pl.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(pl.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pl.PLError, match="no CPU fallback"):
        pl.ORBextractor(1000, 1.2, 8, 20, 7)


def test_product_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "pl-slam_b200", "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".cc")):
            src = open(path, errors="ignore").read()
            assert "import oracle" not in src and "liboracle" not in src and "oracle/" not in src, path
