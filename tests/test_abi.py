"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every symbol that
include/pyrad_b200.h declares, the ctypes table covers the header one to one, and -- with no device --
the product path fails loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pyrad_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from pyrad_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libpyrad_b200.so does not export %s" % s
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes and header differ"
    assert lib.prb_abi_version() == 3


def test_header_has_no_torch_or_cxx_types():
    text = open(os.path.join(ROOT, "include", "pyrad_b200.h")).read()
    assert "torch" not in text.lower() and "std::" not in text and 'extern "C"' in text


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyrad_b200 import _lib
    from pyrad_b200.engine import Engine
    with pytest.raises(_lib.EngineUnavailable):
        Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyrad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
