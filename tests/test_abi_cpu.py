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


def test_every_op_has_a_kernel_reference_and_an_abi_entry():
    """ops.py wrappers <-> oracle/kernel_ref.py twins (same names, compatible signatures) <-> exported C symbols"""
    import inspect
    from oracle import kernel_ref as KR
    from star_b200 import lib as L
    from star_b200 import ops
    skip = {"trace_begin", "trace_end", "launch_count"}
    names = [n for n, f in vars(ops).items() if inspect.isfunction(f) and not n.startswith("_") and n not in skip
             and f.__module__ == ops.__name__]
    assert len(names) >= 20
    src = inspect.getsource(ops)
    for n in names:
        assert hasattr(KR, n), f"oracle/kernel_ref.py has no reference for ops.{n}"
        po = list(inspect.signature(inspect.unwrap(getattr(ops, n))).parameters)
        pr = list(inspect.signature(getattr(KR, n)).parameters)
        n_common = min(len(po), len(pr))
        diff = [i for i in range(n_common) if po[i] != pr[i]]
        assert len(diff) <= 1 and abs(len(po) - len(pr)) <= 1, f"{n}: ops{po} vs kernel_ref{pr}"
    used = {s for s in L.SIGNATURES if f"L.{s}(" in src or f"L.{s} " in src}
    helpers = {"star_version", "star_last_error", "star_init", "star_launch_count"}
    missing = [s for s in L.SIGNATURES if s not in used and s not in helpers]
    assert not missing, f"C entry points without an ops wrapper: {missing}"
