"""CPU-only: the C-ABI library builds, loads, and exports exactly what include/ising_b200.h declares.
No compute call is made here (there is no GPU and no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ising_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(isb_[a-z0-9_]+)\s*\(", txt)))


def test_header_compiles_as_c():
    src = '#include "ising_b200.h"\nint main(void){return ISB_OK;}\n'
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                          "-x", "c", "-"], input=src, text=True, capture_output=True)
    assert res.returncode == 0, res.stderr


def test_library_exports_every_declared_symbol(pkg):
    from isingmodel_jl_b200 import _lib
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ising_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table out of sync with the header"
    assert L.isb_version() == 100


def test_no_oracle_in_product():
    """The product must never import, link or call the oracle (test infrastructure)."""
    pk = os.path.join(ROOT, "isingmodel.jl_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".jl", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "ising_oracle" not in txt, (dp, f)
    out = subprocess.run(["ldd", os.path.join(pk, "libising_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_fails_loudly_without_gpu(pkg):
    """No device -> ISB_ERR_CUDA with a message, never a silent CPU path."""
    from isingmodel_jl_b200 import _lib
    if _lib.load().isb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.IsbError) as ei:
        _lib.Context(0)
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)
    from isingmodel_jl_b200 import SpinSystems
    ss = SpinSystems.SpinSystem(np.array([-1, 1]), np.array([[0.0, 1.0], [1.0, 0.0]]), np.zeros(2))
    with pytest.raises(_lib.IsbError):
        SpinSystems.calcEnergy(ss)
