"""The C-ABI shared library loads and exports every symbol include/feng_b200.h declares (no compute calls: this
runs without a GPU), and a compute entry point fails loudly -- never falls back -- when no device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    with open(os.path.join(ROOT, "include", "feng_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from feng_b200 import build, capi
    build.build()
    names = _header_functions()
    assert len(names) >= 40
    lib = C.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == names


def test_header_cites_the_reference_interfaces():
    with open(os.path.join(ROOT, "include", "feng_b200.h")) as f:
        src = f.read()
    for cite in ("src/feLinearSystem.h", "src/feBilinearForm.cpp", "src/feCompressedRowStorage.cpp",
                 "src/feLinearSystemMklPardiso.cpp", "src/feSysElm.h"):
        assert cite in src


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from feng_b200 import capi
    with pytest.raises(capi.B200Error):
        capi.System(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "feng_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    s = f.read()
                assert not re.search(r"^\s*(import oracle|from oracle)", s, flags=re.M), fn
                assert "libfeng_ref" not in s and "_ref/" not in s, fn
