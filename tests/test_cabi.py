"""The C-ABI library loads and exports every symbol include/pluto_gpu.h
declares (no compute calls: this runs without a GPU)."""
import os
import re

import pytest

from tests.util import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "pluto_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pluto_gpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    from pluto_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    L = _lib.load_library()
    names = _declared()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/pluto_gpu.h but not exported"
    assert sorted(_lib.SYMBOLS) == names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pluto_b200 import GpuStepper
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        GpuStepper(2, (16, 16, 1), (0.1, 0.1))


def test_next_dt_matches_oracle():
    from pluto_b200 import _lib
    from oracle.oracle_lib import next_dt
    L = _lib.load_library()
    for inv, cfl, var, dt in [(3.7, 0.4, 1.1, 1e-3), (120.0, 0.3, 1.1, 5e-2), (0.25, 0.4, 1.1, 1.0)]:
        assert L.pluto_gpu_next_dt(inv, cfl, var, dt) == next_dt(inv, cfl, var, dt)


def test_product_does_not_import_oracle():
    pk = os.path.join(ROOT, "pluto_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
