"""The C-ABI shared library loads without a GPU and exports every symbol include/afterqc_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "afterqc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    # every aqc_* function + the two libed.so-compatible symbols (editdistance/_editdistance.h:16,23)
    return sorted(set(re.findall(r"\b(aqc_[a-z0-9_]+)\s*\(", hdr)) | set(re.findall(r"\b(edit_distance|seek_overlap)\s*\(const char", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from afterqc_b200 import build, _native
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_native.SYMBOLS) == syms, (set(syms) ^ set(_native.SYMBOLS))
    L = _native.lib()
    assert L.aqc_abi_version() == 2


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under afterqc_b200/ may reference it."""
    pkg = os.path.join(ROOT, "afterqc_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "aqc_oracle" not in src and "aqo_" not in src, fn


def test_engine_fails_loudly_without_gpu_or_library(monkeypatch):
    from afterqc_b200 import _abi, _native
    import pytest
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        from afterqc_b200.engine import Engine, EngineError
        with pytest.raises(EngineError):
            Engine(_abi.Params.defaults())
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libafterqc_b200.so")
    with pytest.raises(ImportError):
        _native.lib()


def test_hardware_verified_kernels_are_unchanged():
    """profiles/sass_fingerprints.json lists the kernels whose parity was seen green on a B200; a refactoring of the shared
    headers must leave their SASS untouched (same nvcc), or the file is updated after they are verified on the GPU again:
    `python tools/sass_fingerprint.py --update`."""
    import shutil
    import sys
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_fingerprint
    from afterqc_b200 import build
    build.build()
    assert sass_fingerprint.compare() == []
