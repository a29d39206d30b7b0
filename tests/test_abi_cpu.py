"""CPU: the C-ABI library loads, exports every symbol include/csfm.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "csfm.h")).read()
    return sorted(set(re.findall(r"CSFM_API[^;(]*?\b(csfm_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes

    import csfm_b200
    L = csfm_b200.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/csfm.h but not exported"
    assert set(declared) == set(csfm_b200.SIGNATURES), "binding.py and csfm.h disagree"
    assert b"sm_100a" in L.csfm_version()
    assert isinstance(ctypes.c_char_p(L.csfm_last_error()).value, bytes)


def test_no_cpu_fallback():
    import torch

    import csfm_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(csfm_b200.CsfmError, match="no usable CUDA device"):
        csfm_b200.FMIndex.build_from_text(b"banana$")


def test_product_never_imports_oracle():
    """The product path must not reference oracle/ (only tests/, smoke() and bench.py may)."""
    pkg = os.path.join(ROOT, "compressed-fm-index-implementation-with-learned-optimizations_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".sh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "liboracle" not in src and "libcsref" not in src, f
