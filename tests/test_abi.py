"""CPU: libspg.so loads and exports every symbol include/spg.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

import stark_perpetual_b200 as spg

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "spg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    path = spg.lib_path()
    assert os.path.exists(path), "libspg.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 8
    for name in names:
        assert hasattr(lib, name), "missing export: " + name


def declared_prototypes():
    """name -> number of parameters, from include/spg.h"""
    text = open(os.path.join(ROOT, "include", "spg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(spg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_python_binding_declares_every_prototype():
    """Every entry point has ctypes argtypes of the header's arity in the Python binding: an undeclared size_t passed
    as a Python int is truncated to 32 bits and the upper half of the stack slot is garbage."""
    from stark_perpetual_b200 import _lib
    lib = _lib._load()
    protos = declared_prototypes()
    assert set(protos) == set(declared_symbols())
    for name, n_params in protos.items():
        fn = getattr(lib, name)
        assert fn.argtypes is not None, "no argtypes for " + name
        assert len(fn.argtypes) == n_params, (name, len(fn.argtypes), n_params)


def test_no_cpu_fallback():
    """Without a GPU the product path must refuse to run rather than fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(spg.SpgError):
        spg.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "stark_perpetual_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "spg_oracle" not in src, f
