"""The C-ABI shared library: it loads, exports every symbol include/axctd.h
declares, and refuses to work without a GPU.  CPU only (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from axctdprocessor_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "axctd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(axctd_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), n
    from axctdprocessor_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names


def test_struct_sizes_agree(lib):
    from axctdprocessor_b200 import _lib
    for which, st in enumerate((_lib.ConfigDesc, _lib.DropSummary, _lib.Frame, _lib.Chunk)):
        assert lib.axctd_struct_size(which) == C.sizeof(st)
    assert lib.axctd_abi_version() == _lib.ABI_VERSION == 4 and lib.axctd_has_cuda() == 1


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from axctdprocessor_b200 import engine
    with pytest.raises(RuntimeError):
        engine.Engine(0)


def test_product_refuses_emulation_library():
    from emu_util import build_emu
    from axctdprocessor_b200 import _lib, engine
    emu = _lib.bind(C.CDLL(build_emu()))
    assert emu.axctd_has_cuda() == 0
    with pytest.raises(RuntimeError):
        engine.Engine(lib=emu)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "axctdprocessor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "axctd_oracle" not in src and "ref_shim" not in src, f
