"""numpy twin of the C oracle — TEST INFRASTRUCTURE ONLY (second opinion with a different FFT).

Restates caf_rust's arithmetic with numpy's pocketfft so that a mistake in the hand-written
Stockham FFT of caf_oracle.c cannot hide: the two must agree to ~1e-15 of the surface maximum.
File:line citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np


def apply_freq_shift(samples, freq_shift: float, fs: int) -> np.ndarray:
    """mod.rs:46-65, recursion kept (acc *= shift) so the rounding pattern is the reference's."""
    x = np.asarray(samples, dtype=np.complex128)
    dt = 1.0 / float(fs)
    theta = 2.0 * np.pi * freq_shift * dt
    shift = complex(np.cos(theta), np.sin(theta))
    acc = np.empty(x.size, dtype=np.complex128)
    a = 1.0 + 0.0j
    for i in range(x.size):  # small inputs only; the C oracle is the fast one
        acc[i] = a
        a = complex(a.real * shift.real - a.imag * shift.imag, a.real * shift.imag + a.imag * shift.real)
    return x * acc


def apply_freq_shift_direct(samples, freq_shift: float, fs: int) -> np.ndarray:
    """Closed form of the same shift, exp(+j 2 pi f n / fs) per sample (no error accumulation)."""
    x = np.asarray(samples, dtype=np.complex128)
    n = np.arange(x.size, dtype=np.float64)
    return x * np.exp(2j * np.pi * (freq_shift / float(fs)) * n)


def xcor(a, b) -> np.ndarray:
    """xcor_rustfft.rs:51-78: IFFT_unnormalised( FFT(a) * conj(FFT(b)) / n )."""
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    assert a.size == b.size
    n = a.size
    prod = np.fft.fft(a) * np.conj(np.fft.fft(b)) / n
    return np.fft.ifft(prod) * n  # numpy's ifft divides by n; the reference's inverse does not


def caf_surface(needle, haystack, freqs_hz, fs: int, direct_phasor: bool = True):
    """mod.rs:121-166.  Returns (surface[D,2L], row_peak_idx, row_peak_val)."""
    needle = np.asarray(needle, dtype=np.complex128)
    haystack = np.asarray(haystack, dtype=np.complex128)
    assert needle.size == haystack.size
    l = needle.size
    npad = np.concatenate([needle, np.zeros(l, dtype=np.complex128)])     # :130
    hpad = np.concatenate([haystack, np.zeros(l, dtype=np.complex128)])   # :131
    fh = np.fft.fft(hpad)
    d = len(freqs_hz)
    surf = np.empty((d, 2 * l), dtype=np.float64)
    shift = apply_freq_shift_direct if direct_phasor else apply_freq_shift
    for r, f in enumerate(freqs_hz):
        sh = shift(npad, float(f), fs)                                    # :138
        res = np.fft.ifft(fh * np.conj(np.fft.fft(sh)))                   # :139 (1/n folded)
        surf[r] = res.real * res.real + res.imag * res.imag               # norm_sqr :147
    pidx = np.argmax(surf, axis=1).astype(np.uint64)                      # first max == strict >
    pval = surf[np.arange(d), pidx] if d else np.zeros(0)
    return surf, pidx, pval


def find_peak(freqs_hz, row_peak_idx, row_peak_val):
    """mod.rs:31-42."""
    best, f, idx = 0.0, 0.0, 0
    for r in range(len(freqs_hz)):
        if row_peak_val[r] > best:
            best, f, idx = row_peak_val[r], float(freqs_hz[r]), int(row_peak_idx[r])
    return f, idx
