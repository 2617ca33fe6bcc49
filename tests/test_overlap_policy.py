"""CPU: the host-side bookkeeping of caf_b200_set_overlap (caf_cookoff_b200/csrc/overlap_policy.hpp) -- which launches
may skip the wait for the grid before them -- compiled with g++ and run on its own: read-after-write, write-after-read
and write-after-write against every launch of the history, shared read-only inputs, the history restarting after a launch
that waited or after any other library launch, aliases older than the history, empty ranges, the grid of each mode."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_overlap_policy(tmp_path):
    gxx = shutil.which("g++")
    assert gxx, "g++ is required"
    exe = str(tmp_path / "test_overlap_policy")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_overlap_policy.cpp"), "-o", exe])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all checks passed" in res.stdout
