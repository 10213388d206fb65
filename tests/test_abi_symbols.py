"""The C-ABI library must load and export every symbol include/fers_b200.h declares (no compute)."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "fers_b200.h")).read()
    return sorted(set(re.findall(r"\b(fers_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from fers_core_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fers_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)


def test_version_and_loud_failure_without_gpu():
    import pytest
    import torch
    from fers_core_b200 import _lib, engine
    lib = _lib.load()
    assert b"sm_100a" in lib.fers_b200_version()
    if not torch.cuda.is_available():
        with pytest.raises(engine.EngineError):
            engine.calculate_from_json('{"model": {"nodes": []}, "analysis": {}}')


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for pkg in ("fers_core_b200", "fers_calculations"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
