"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/oak_b200.h declares; host-only entry points work; GPU entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from oak_b200 import build
    build.build()
    from oak_b200 import _lib
    return _lib.lib()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "oak_b200.h")).read()
    return sorted(set(re.findall(r"OAKB200_API[^;]*?\b(oakb200_\w+)\s*\(", txt)))


def test_header_symbols_all_exported(lib):
    from oak_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/oak_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)


def test_library_is_sm100a_only():
    import subprocess
    from oak_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_partition_matches_parallpartion(lib):
    # parall.F90:176-177: start = nzones*cum(p)/total + 1 ; end = nzones*cum(p+1)/total (integer division)
    import oak_b200
    from oak_b200 import dist
    for nz, P in [(10, 4), (1000000, 8), (7, 3), (3, 8), (0, 2)]:
        first = oak_b200.partition_zones(nz, P)
        want = [(nz * p) // P for p in range(P + 1)]
        assert list(first) == want
        assert list(dist.partition(nz, P)) == want
        assert first[0] == 0 and first[-1] == nz and (np.diff(first) >= 0).all()


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import oak_b200
    with pytest.raises(oak_b200.OakB200Error) as e:
        oak_b200.Handle(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under oak_b200/ or include/ may reference it
    for base in ("oak_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(import|from)\s+oracle|oak_oracle|liboak_oracle", txt, re.M), (dp, f)
