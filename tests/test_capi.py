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


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/plslam_b200.h compiles as C99 (no C++ types in the signatures) and every exported
    pl_* symbol of the library is declared in it."""
    import subprocess
    src = tmp_path / "cabi.c"
    src.write_text('#include "plslam_b200.h"\nint main(void) { return pl_version() < 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "cabi.o")])
    out = subprocess.check_output(["nm", "-D", "--defined-only", pl.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r"\bT (pl_[a-z0-9_]+)\b", out)))
    undeclared = [s for s in exported if s not in declared_symbols()]
    assert not undeclared, undeclared
