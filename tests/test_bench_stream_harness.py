"""CPU check of the second timing harness (scripts/bench_stream.measure, which bench.py also calls at N = 1): the
rotation over seeded pairs and surface buffers, the event bracket, the launch count and the planted-peak check, run
against a stand-in library whose `caf_b200_batch_f64_dev` answers from the CPU oracle.  What is under test is the
harness, not the kernels (those are measured on the GPU); torch's CUDA events are replaced by wall-clock stand-ins."""
import ctypes as C
import os
import sys
import time

import numpy as np
import pytest

from conftest import ROOT


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _FakeLib:
    """The four entry points measure() uses, over host memory: pointers are read / written with ctypes."""

    def __init__(self, break_pair=None):
        self.launches = 0
        self.break_pair = break_pair
        self.seen = []

    def caf_b200_batch_f64_dev(self, h, n_ptr, h_ptr, p, l, f_ptr, d, fs, surf_ptr, rv_ptr, ri_ptr, pk_ptr):
        from oracle import oracle as O
        assert p == 1 and l == 4096 and d == 400 and fs == 48000 and surf_ptr and rv_ptr and ri_ptr and pk_ptr
        needle = np.ctypeslib.as_array(C.cast(n_ptr, C.POINTER(C.c_double)), (2 * l,)).view(np.complex128)
        hay = np.ctypeslib.as_array(C.cast(h_ptr, C.POINTER(C.c_double)), (2 * l,)).view(np.complex128)
        freqs = np.ctypeslib.as_array(C.cast(f_ptr, C.POINTER(C.c_double)), (d,))
        _, pidx, pval = O.caf_surface(needle.copy(), hay.copy(), freqs.copy(), fs, want_surface=False, threads=4)
        f, lag = O.find_peak(freqs, pidx, pval)
        if self.break_pair is not None and len(self.seen) % 6 == self.break_pair:
            lag += 1
        out = np.ctypeslib.as_array(C.cast(pk_ptr, C.POINTER(C.c_uint64)), (4,))
        out[0:1] = np.array([pval.max()]).view(np.uint64)
        out[1:2] = np.array([f]).view(np.uint64)
        out[2], out[3] = int(np.argmax(pval)), lag
        self.seen.append((n_ptr, surf_ptr))
        self.launches += 1
        return 0

    def caf_b200_set_overlap(self, h, on):
        self.overlap_calls = getattr(self, "overlap_calls", []) + [int(on)]
        return 0

    def caf_b200_launch_count(self, h):
        return self.launches

    def caf_b200_probe_fma_tflops(self, h, is_f64, out):
        out._obj.value = 37.0
        return 0

    def caf_b200_last_error(self):
        return b""


class _Handle:
    raw = C.c_void_p(1)


@pytest.fixture
def bench_stream(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import bench_stream as bs
    yield bs
    sys.path.remove(os.path.join(ROOT, "scripts"))


def test_rotation_launch_count_and_planted_peak_check(bench_stream):
    import torch
    lib = _FakeLib()
    res = bench_stream.measure(lib, _Handle(), None, torch.device("cpu"), pairs=6, surfaces=4, steps=9, warmup=3)
    assert res["gpu_launches"] == 9 and lib.launches == 12 and res["steps"] == 9
    assert res["pairs_checked"] == 6 and res["peaks_off"] == []
    assert res["us_per_surface"] > 0 and abs(res["cells_per_s"] - 400 * 8192 / (res["us_per_surface"] * 1e-6)) < 1e-3 * res["cells_per_s"]
    assert abs(res["frac"] - res["tflops"] / 37.0) < 1e-12
    assert abs(res["working_set_mb"] - (6 * 2 * 4096 * 16 + 4 * 400 * 8192 * 8) / 1e6) < 1e-9
    # launch k uses pair k mod 6 and surface buffer k mod 4
    assert len({s[0] for s in lib.seen}) == 6 and len({s[1] for s in lib.seen}) == 4
    assert [s[0] for s in lib.seen[:6]] == [s[0] for s in lib.seen[6:12]]
    assert lib.seen[0][1] == lib.seen[4][1] != lib.seen[1][1]


def test_a_wrong_peak_is_reported(bench_stream):
    import torch
    lib = _FakeLib(break_pair=2)
    res = bench_stream.measure(lib, _Handle(), None, torch.device("cpu"), pairs=6, surfaces=2, steps=6, warmup=0)
    assert [b[0] for b in res["peaks_off"]] == [2]
