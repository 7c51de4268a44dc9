"""The C-ABI library loads on a CPU-only box and exports every symbol include/star_sm100.h declares
(no compute calls here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "star_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(star_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header():
    from star_b200 import lib
    L = lib.get_lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/star_sm100.h but not exported"
    assert sorted(lib.SIGNATURES) == syms, "star_b200/lib.py SIGNATURES and the header disagree"
    assert L.star_version() == 100


def test_init_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from star_b200 import lib
    with pytest.raises(lib.StarError):
        lib.ensure_init(0)
