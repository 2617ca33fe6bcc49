"""The C++ host mirror (include/caf_b200.hpp) restating caf_rust/tests/test.rs, compiled with g++ against the in-tree
libcaf_b200.so and executed on the GPU."""
import os
import shutil
import subprocess

import pytest

from caf_cookoff_b200 import _lib
from conftest import DATA, ROOT

BIN = os.path.join(ROOT, "tests", "cpp", "test_rs")


def _build():
    gxx = shutil.which("g++")
    assert gxx, "g++ is required for the C++ mirror test"
    so_dir = os.path.dirname(_lib.SO_PATH)
    cmd = [gxx, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_rs.cpp"),
           "-o", BIN, "-L", so_dir, "-lcaf_b200", f"-Wl,-rpath,{so_dir}"]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_without_gpu():
    """CPU check: the header-only mirror and the C ABI header compile and link against the built library."""
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_restatement_of_test_rs():
    _build()
    res = subprocess.run([BIN, DATA], capture_output=True, text=True, timeout=300)
    print(res.stdout[-3000:], res.stderr[-2000:])
    assert res.returncode == 0
    assert "all tests passed" in res.stdout
    assert "Frequency offset: 69.0Hz" in res.stdout and "Time offset: 202 samples (4.208ms)" in res.stdout
