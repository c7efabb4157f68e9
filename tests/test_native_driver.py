"""tests/native/c_driver.c: a plain-C program (gcc, no Python / torch / CUDA headers) that feeds the C ABI with
row blocks in the reference's HPCSparseMatrix storage (1-based, compressed columns) and checks the assembled
gradient / R'HR against a dense evaluation written in C."""
import os
import subprocess

import pytest

from mgb_b200 import build as _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = str(tmp_path / "c_driver")
    libdir = os.path.dirname(_build.LIB)
    cmd = ["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tests", "native", "c_driver.c"), "-L" + libdir, "-lmgb_b200", "-lm",
           "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_driver_symbolic(tmp_path):
    from mgb_b200 import capi
    capi.load()
    res = subprocess.run([_compile(tmp_path), "--symbolic"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "C_DRIVER_OK" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_c_driver_numeric(tmp_path):
    from mgb_b200 import capi
    capi.load()
    res = subprocess.run([_compile(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "C_DRIVER_OK" in res.stdout, res.stdout + res.stderr
