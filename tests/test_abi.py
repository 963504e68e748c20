"""The C-ABI shared library loads and exports every symbol include/femb200.h declares; without a GPU
every compute entry fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "femb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(femb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from fempy_b200 import _lib

    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in femb200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes signatures out of sync with the header"
    assert b"sm_100a" in lib.femb_version()


def test_material_struct_layout():
    from fempy_b200._lib import Material

    assert C.sizeof(Material) == 32
    assert Material.model.offset == 24 and Material.nonlinear.offset == 28


def test_binary_is_sm100a_only():
    """the shipped cubin targets sm_100a and nothing else (no multi-arch fallback)"""
    import subprocess

    from fempy_b200._build import LIB_PATH

    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(lib):
    """Without a CUDA device plan creation must fail with a clear error; with one it must succeed."""
    from fempy_b200 import AssemblyPlan
    from fempy_b200._lib import FembError

    if lib.femb_device_count() == 0:
        with pytest.raises(FembError, match="no CUDA device"):
            AssemblyPlan(10)
        from fempy_b200 import AssemblyUtils

        with pytest.raises(FembError, match="no CUDA device"):
            AssemblyUtils.localMatricesToCOOArrays(np.ones((1, 2, 2)), np.array([[0, 1]]))
    else:
        AssemblyPlan(10).close()


def test_package_never_imports_the_oracle():
    """product code must not reach into oracle/ (test infrastructure)"""
    pkg = os.path.join(ROOT, "fempy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "fem_oracle" not in text, f
