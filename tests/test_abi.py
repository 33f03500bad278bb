"""The product library loads on a machine without a GPU and exports every symbol include/cp3b200.h declares;
without a device the hot path fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import riss_solver_b200 as rs
from harness import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cp3b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(CP[A-Za-z]+|cp3g_[a-z_0-9]+)\s*\(", src)) - {"CP3G_"})


def test_library_exports_every_declared_symbol():
    if not os.path.exists(rs.DEFAULT_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "riss-solver_b200", "csrc")])
    L = C.CDLL(rs.DEFAULT_LIB)
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), s
    assert set(rs.CP_SYMBOLS + rs.CP3G_SYMBOLS) <= set(syms)


def test_product_has_no_cpu_path():
    """no compute call without a device: the library must refuse, loudly"""
    L = rs.load()
    if L.cp3g_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert not L.cp3g_create(0)
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; import riss_solver_b200 as rs;"
            "cp = rs.Coprocessor(['-enabled_cp3','-subsimp']); cp.add_stream(np.array([1,2,0,1,2,3,0],np.int32)); cp.preprocess(); print('SURVIVED')" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode != 0 and "SURVIVED" not in p.stdout
    assert "no CPU fallback" in p.stderr


def test_product_library_does_not_link_the_oracle_or_mock():
    out = subprocess.run(["nm", "-D", "--defined-only", rs.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "cp3o_" not in out            # oracle symbols
    src = open(os.path.join(ROOT, "riss-solver_b200", "csrc", "Makefile")).read()
    assert "oracle" not in src and "mock" not in src
