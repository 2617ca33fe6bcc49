"""The reference's Go and Python programs behind the same GPU path.

caf_cookoff holds three programs that compute one filterbank CAF with three conventions.  The Rust crate is the
oracle of this repository; this module reproduces what the other two return, from the same kernels, through
`caf_b200_surface_layout_*` (the conversion runs on the GPU):

    Python  /root/reference/caf_python/caf.py
        apply_fdoa(ray, fdoa, samp_rate)                 caf.py:28-33    == apply_freq_shift (same sign)
        amb_surf(needle, haystack, freqs_hz, samp_rate)  caf.py:101-122  -> [D, L] |correlate(shifted, haystack, 'same')|
        tau_max = len(needle)//2 - tmax                  caf.py:144-146
    Go      /root/reference/caf_go/caf.go, main.go
        apply_fdoa(ray, fdoa, samp_rate)                 caf.go:118-126
        amb_surf(needle, haystack, freqs_hz, samp_rate)  caf.go:162-173  -> [D][2L] |IFFT(FFT(a|0) conj(FFT(0|b)))|
        find_2d_peak(surf) -> (fdx, tdx, max)            caf.go:183-195  first strict-> maximum, row-major
        main.go:35 reports len(apple) - tdx samples
        dump_surf(path, surf)                            caf.go:14-29    row-major little-endian float64

Magnitudes are sqrt(re^2 + im^2) of the fp64 correlation (np.abs / cmplx.Abs use hypot: last-ulp differences).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .api import CafB200, CafPanic, Handle, _check, _ptr, default_handle

LAYOUT_RUST, LAYOUT_PYTHON, LAYOUT_GO = 0, 1, 2


def _whole_sample_rate(samp_rate) -> int:
    """The siblings take a float sample rate (caf.py:28, caf.go:118); the C ABI takes the Rust crate's u32 (mod.rs:46).
    A rate the u32 cannot hold exactly would silently give a different phasor than the sibling program: refuse it."""
    fs = float(samp_rate)
    if not (fs >= 1.0 and fs <= 4294967295.0 and fs == int(fs)):
        raise ValueError(f"samp_rate must be a whole number of Hz in [1, 2^32 - 1] (got {samp_rate!r}): the library's "
                         "sample rate is the Rust crate's u32 (mod.rs:46)")
    return int(fs)


def surface_layout(needle, haystack, freqs_hz, fs, layout: int, *, f32: bool = False, want_surface: bool = True,
                   handle: Optional[Handle] = None):
    """caf_b200_surface_layout_{f64,f32}: returns (surface [D, W] or None, Peak) in the sibling's layout."""
    cdt, rdt, sfx = (np.complex64, np.float32, "f32") if f32 else (np.complex128, np.float64, "f64")
    fs = _whole_sample_rate(fs)
    n_ = np.ascontiguousarray(needle, dtype=cdt).ravel()
    h_ = np.ascontiguousarray(haystack, dtype=cdt).ravel()
    f_ = np.ascontiguousarray(freqs_hz, dtype=np.float64).ravel()
    if n_.size != h_.size:
        raise CafPanic("assert len_needle == len_haystack (caf.py:41,107) / panic(\"input arrays should be same size\") (caf.go:97-99)")
    l, d = n_.size, f_.size
    w = l if layout == LAYOUT_PYTHON else 2 * l
    out = np.empty((d, w), dtype=rdt) if want_surface else None
    pk = _lib.Peak()
    h = handle or default_handle()
    fn = getattr(_lib.load(), "caf_b200_surface_layout_" + sfx)
    _check(fn(h.raw, _ptr(n_), _ptr(h_), l, _ptr(f_), d, fs, int(layout), _ptr(out),
              C.cast(C.byref(pk), C.c_void_p)))
    return out, pk


class PythonSibling:
    """caf_python/caf.py: same function names, argument order and return shapes."""

    @staticmethod
    def apply_fdoa(ray, fdoa: float, samp_rate: float) -> np.ndarray:
        return CafB200.apply_freq_shift(ray, float(fdoa), _whole_sample_rate(samp_rate))

    @staticmethod
    def amb_surf(needle, haystack, freqs_hz, samp_rate: float) -> np.ndarray:
        return surface_layout(needle, haystack, freqs_hz, samp_rate, LAYOUT_PYTHON)[0]

    # the four variants of caf.py differ only in how they schedule rows on the CPU (caf.py:35-99)
    amb_surf_numba = amb_surf_multiprocessing = amb_surf_multiprocessing_numba = amb_surf

    @staticmethod
    def peak(needle, haystack, freqs_hz, samp_rate: float) -> Tuple[int, float]:
        """(tau_max, freq_max) as caf.py:144-146 prints them; the surface stays on the GPU."""
        _, pk = surface_layout(needle, haystack, freqs_hz, samp_rate, LAYOUT_PYTHON, want_surface=False)
        if pk.doppler_idx == (1 << 64) - 1:          # np.argmax of an all-zero surface is (0, 0)
            return len(np.atleast_1d(needle)) // 2, float(np.atleast_1d(freqs_hz)[0])
        return len(np.atleast_1d(needle)) // 2 - int(pk.delay_idx), float(pk.freq_hz)


class GoSibling:
    """caf_go/caf.go + main.go."""

    @staticmethod
    def apply_fdoa(ray, fdoa: float, samp_rate: float) -> np.ndarray:
        return CafB200.apply_freq_shift(ray, float(fdoa), _whole_sample_rate(samp_rate))

    @staticmethod
    def amb_surf(needle, haystack, freqs_hz, samp_rate: float) -> np.ndarray:
        return surface_layout(needle, haystack, freqs_hz, samp_rate, LAYOUT_GO)[0]

    amb_surf_concurrent = amb_surf

    @staticmethod
    def find_2d_peak(needle, haystack, freqs_hz, samp_rate: float) -> Tuple[int, int, float]:
        """(fdx, tdx, max) of caf.go:183-195 without moving the surface to the host; (0, 0, 0.0) when nothing is > 0."""
        _, pk = surface_layout(needle, haystack, freqs_hz, samp_rate, LAYOUT_GO, want_surface=False)
        if pk.doppler_idx == (1 << 64) - 1:
            return 0, 0, 0.0
        return int(pk.doppler_idx), int(pk.delay_idx), float(pk.value)

    @staticmethod
    def dump_surf(path: str, surf) -> None:
        np.ascontiguousarray(surf, dtype="<f8").tofile(path)


def arange(start: float, stop: float, step: float) -> np.ndarray:
    """caf.go:175-181: repeated addition (not numpy's start + i*step), so the grid carries Go's rounding."""
    out, x = [], float(start)
    while x < stop:
        out.append(x)
        x += step
    return np.array(out, dtype=np.float64)
