"""ctypes wrapper over oracle/libcaf_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (caf_cookoff_b200/) never does.

Every function restates a reference site; see oracle/caf_oracle.c for file:line citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcaf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "caf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcaf_oracle.so", "CC=gcc"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u32, dbl = C.c_void_p, C.c_size_t, C.c_uint32, C.c_double
        L.oracle_apply_freq_shift.argtypes = [vp, sz, dbl, u32, vp]
        L.oracle_apply_freq_shift.restype = None
        L.oracle_xcor_run.argtypes = [vp, sz, vp, sz, vp]
        L.oracle_xcor_run.restype = C.c_int
        L.oracle_caf_surface.argtypes = [vp, sz, vp, sz, vp, sz, u32, vp, vp, vp]
        L.oracle_caf_surface.restype = C.c_int
        L.oracle_caf_surface_threadpool.argtypes = [vp, sz, vp, sz, vp, sz, u32, C.c_int, vp, vp, vp]
        L.oracle_caf_surface_threadpool.restype = C.c_int
        L.oracle_find_peak.argtypes = [vp, vp, vp, sz, C.POINTER(dbl), C.POINTER(C.c_uint64)]
        L.oracle_find_peak.restype = None
        L.oracle_direct_cells_ld.argtypes = [vp, vp, sz, dbl, u32, vp, sz, vp]
        L.oracle_direct_cells_ld.restype = None
        L.oracle_widen_c64.argtypes = [vp, sz, vp]
        L.oracle_widen_c64.restype = None
        _lib = L
    return _lib


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def apply_freq_shift(samples, freq_shift: float, fs: int) -> np.ndarray:
    """caf_rust/src/caf/mod.rs:46-65."""
    x = _c128(samples)
    out = np.empty_like(x)
    lib().oracle_apply_freq_shift(_p(x), x.size, float(freq_shift), int(fs), _p(out))
    return out


def xcor(a, b) -> np.ndarray:
    """Xcor::run(a, b) — caf_rust/src/caf/xcor_rustfft.rs:51-78.  AssertionError mirrors :54-55."""
    a, b = _c128(a), _c128(b)
    out = np.empty_like(a)
    if lib().oracle_xcor_run(_p(a), a.size, _p(b), b.size, _p(out)) != 0:
        raise AssertionError("a.len() == self.n / b.len() == self.n (xcor_rustfft.rs:54-55)")
    return out


def caf_surface(needle, haystack, freqs_hz, fs: int, *, want_surface: bool = True, threads: int = 0):
    """CafRustFFT::caf_surface (mod.rs:121-166); threads>0 -> CafRustFFTThreadpool (mod.rs:391-461).

    Returns (surface[D,2L] or None, row_peak_idx[D] uint64, row_peak_val[D] float64).
    """
    n_, h_ = _c128(needle), _c128(haystack)
    f_ = np.ascontiguousarray(freqs_hz, dtype=np.float64)
    d, n = f_.size, 2 * n_.size
    surf = np.empty((d, n), dtype=np.float64) if want_surface else None
    pidx = np.zeros(d, dtype=np.uint64)
    pval = np.zeros(d, dtype=np.float64)
    sp = _p(surf) if surf is not None else None
    if threads and threads > 0:
        rc = lib().oracle_caf_surface_threadpool(_p(n_), n_.size, _p(h_), h_.size, _p(f_), d, int(fs),
                                                 int(threads), sp, _p(pidx), _p(pval))
    else:
        rc = lib().oracle_caf_surface(_p(n_), n_.size, _p(h_), h_.size, _p(f_), d, int(fs),
                                      sp, _p(pidx), _p(pval))
    if rc != 0:
        raise AssertionError("needle/haystack length mismatch (xcor_rustfft.rs:54-55)")
    return surf, pidx, pval


def find_peak(freqs_hz, row_peak_idx, row_peak_val):
    """CafSurface::find_peak — mod.rs:31-42.  Returns (freq, samp_idx)."""
    f_ = np.ascontiguousarray(freqs_hz, dtype=np.float64)
    i_ = np.ascontiguousarray(row_peak_idx, dtype=np.uint64)
    v_ = np.ascontiguousarray(row_peak_val, dtype=np.float64)
    fo, io = C.c_double(0.0), C.c_uint64(0)
    lib().oracle_find_peak(_p(f_), _p(i_), _p(v_), f_.size, C.byref(fo), C.byref(io))
    return fo.value, int(io.value)


def direct_cells(needle, haystack, freq: float, fs: int, lags) -> np.ndarray:
    """Long-double direct evaluation of selected surface cells (no FFT, no phasor recursion)."""
    n_, h_ = _c128(needle), _c128(haystack)
    assert n_.size == h_.size
    l_ = np.ascontiguousarray(lags, dtype=np.uint64)
    out = np.empty(l_.size, dtype=np.float64)
    lib().oracle_direct_cells_ld(_p(n_), _p(h_), n_.size, float(freq), int(fs), _p(l_), l_.size, _p(out))
    return out


def read_file_c64(path: str) -> np.ndarray:
    """caf_rust/src/utils.rs:10-35: packed LE f32 I/Q -> complex128."""
    raw = np.fromfile(path, dtype="<f4")
    nsamp = raw.size // 2
    out = np.empty(nsamp, dtype=np.complex128)
    lib().oracle_widen_c64(_p(np.ascontiguousarray(raw)), nsamp, _p(out))
    return out


def gen_float_shifts(start: float, end: float, step: float) -> np.ndarray:
    """caf_rust/tests/test.rs:335-352 — integer milli-Hz half-open range, /1e3."""
    s, e, st = int(start * 1000.0), int(end * 1000.0), int(step * 1000.0)
    return np.array([m / 1e3 for m in range(s, e, st)], dtype=np.float64)


def bench_shifts() -> np.ndarray:
    """caf_rust/benches/caf_bench.rs:30-35 / src/main.rs:19-22 — 400 rows, -100.0 .. 99.5 Hz."""
    return np.array([m / 1e3 for m in range(-100000, 100000, 500)], dtype=np.float64)
