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


def test_binding_structs_match_the_header(tmp_path):
    """The ctypes mirrors in binding.py have the size and field offsets a C compiler gives the structs of include/csfm.h
    (a field added to one side only would shift everything behind it silently)."""
    import ctypes
    import subprocess

    import csfm_b200
    from csfm_b200 import binding
    pairs = {"csfm_params": binding.Params, "csfm_index_info": binding.IndexInfo, "csfm_call_stats": binding.CallStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "csfm.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in filter(None, out):
        cname, what, val = ln.split()
        cls = pairs[cname]
        want = ctypes.sizeof(cls) if what == "size" else getattr(cls, what).offset
        assert int(val) == want, ln
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
