"""Seeded input generator — a port of the reference's utils/generate.py (file:line cited per step).

Produces the s0/s1 pairs every test and benchmark of caf_rust reads from ../data: band-limited complex
noise "chirps" (s0, `chirp_{i}_raw.c64`) and a delayed, doppler-shifted, noisy copy (s1,
`chirp_{i}_T{lag:+d}samp_F{f:+.2f}Hz.c64`), stored as numpy complex64.  With seed 0 and the default
arguments the byte stream is identical to what the reference script writes (checked in
tests/test_generate.py against tests/golden/data, which came from the unmodified script), because the
legacy MT19937 stream is consumed in the same order:
    seed -> chirp_order, relative_bandwidth, sweep_range        generate.py:42,47-49
    per index: lag (:55), [srange, re, im] inside generate_chirp (:24-25), foffset (:61), noise re/im (:66)
`chirp_length` is a parameter so the larger BASELINE configs (32768, 2**19 samples) use the same recipe.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterator, Optional

import numpy as np
import scipy.signal


@dataclass
class ChirpPair:
    index: int
    raw: np.ndarray        # s0, complex64, chirp_length samples
    search: np.ndarray     # s1, complex64, lag + chirp_length + 96 samples
    lag: int
    foffset_hz: float

    @property
    def raw_name(self) -> str:
        return "chirp_{:d}_raw.c64".format(self.index)

    @property
    def search_name(self) -> str:
        return "chirp_{:d}_T{:+d}samp_F{:+.2f}Hz.c64".format(self.index, self.lag, self.foffset_hz)


def _constant_offset(signal: np.ndarray, dfc: float, sample_rate: float) -> np.ndarray:
    # generate.py:15-16 (scalar branch)
    return np.exp(1j * 2 * np.pi * dfc * np.arange(len(signal)) / sample_rate) * signal


def _varying_offset(signal: np.ndarray, dfc: np.ndarray, sample_rate: float) -> np.ndarray:
    # generate.py:18-19 (array branch) — note the reference adds arange/sample_rate, not a 2*pi term
    phi = np.cumsum(2 * np.pi * dfc) / sample_rate
    return np.exp(1j * (np.arange(len(signal)) / sample_rate + phi)) * signal


def _shaped_noise(rs: np.random.RandomState, sample_rate: float, chirp_length: int, chirp_order: int,
                  relative_bandwidth: float, sweep_range_hz: float) -> np.ndarray:
    # generate.py:21-39
    kernel = scipy.signal.firwin(127, cutoff=0.5 * relative_bandwidth, fs=sample_rate)
    rs.uniform(1e3, 10e3)                                   # `srange`: drawn and unused (:24)
    noise = rs.normal(0, 1, chirp_length) + 1j * rs.normal(0, 1, chirp_length)
    chirp = scipy.signal.filtfilt(kernel, 1, noise)
    chirp = np.hanning(chirp_length) * chirp                # taper (:30-31)
    chirp = chirp.astype(np.complex64)                      # (:32)
    shape = np.linspace(-1, 1, chirp_length) ** chirp_order
    return _varying_offset(chirp, shape * sweep_range_hz, sample_rate)


def pairs(seed: int = 0, count: int = 10, chirp_length: int = 4096, sample_rate: float = 48e3,
          dfc_range_hz: float = 1e2, tail_zeros: int = 96) -> Iterator[ChirpPair]:
    """generate.py:41-68 as a generator of ChirpPair."""
    rs = np.random.RandomState(seed)                        # == np.random.seed(seed) on the global stream
    chirp_order = rs.randint(2, 5)
    relative_bandwidth = rs.uniform(1e-3, 5e-2)
    sweep_range_hz = rs.uniform(1e3, 10e3)
    for idx in range(count):
        lag = int(rs.randint(7, 256))
        raw = _shaped_noise(rs, sample_rate, chirp_length, chirp_order, relative_bandwidth,
                            sweep_range_hz).astype(np.complex64)
        foffset = float(rs.uniform(-dfc_range_hz, dfc_range_hz))
        search = np.concatenate([np.zeros(lag), raw, np.zeros(tail_zeros)])
        search = _constant_offset(search, foffset, sample_rate)
        search = search + (rs.normal(0, 1e-5, len(search)) + 1j * rs.normal(0, 1e-5, len(search)))
        yield ChirpPair(idx, raw, search.astype(np.complex64), lag, foffset)


def write_pairs(data_dir: str, **kw) -> list:
    os.makedirs(data_dir, exist_ok=True)
    names = []
    for p in pairs(**kw):
        p.raw.tofile(os.path.join(data_dir, p.raw_name))
        p.search.tofile(os.path.join(data_dir, p.search_name))
        names += [p.raw_name, p.search_name]
    return names


def pair(index: int = 0, seed: int = 0, chirp_length: int = 4096, **kw) -> ChirpPair:
    """The index-th pair of a seed's stream (index 0, seed 0 = the README benchmark inputs)."""
    for p in pairs(seed=seed, count=index + 1, chirp_length=chirp_length, **kw):
        if p.index == index:
            return p
    raise IndexError(index)


def as_inputs(p: ChirpPair, dtype=np.complex128):
    """(needle, haystack) the way every caller in the reference prepares them: widen to complex128
    (utils.rs:10-35) and cut/zero-extend the haystack to the needle's length (test.rs:19, caf_bench.rs:28)."""
    needle = p.raw.astype(dtype)
    hay = p.search[: needle.size].astype(dtype)
    if hay.size < needle.size:
        hay = np.concatenate([hay, np.zeros(needle.size - hay.size, dtype=dtype)])
    return needle, hay


if __name__ == "__main__":
    import sys
    out = sys.argv[1] if len(sys.argv) > 1 else "../data"
    for n in write_pairs(out):
        print(n)
