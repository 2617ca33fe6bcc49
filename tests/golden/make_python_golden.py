#!/usr/bin/env python
"""Golden surface / shift / correlation VALUES produced by the reference itself.

Run in the build container (where /root/reference is mounted):  python tests/golden/make_python_golden.py

The Rust crate cannot be compiled here (no cargo), and its own tests only pin peak indices.  The reference's Python
sibling, /root/reference/caf_python/caf.py, computes the same filterbank CAF with scipy and CAN be imported, so its
UNMODIFIED functions (`apply_fdoa`, `xcor`, `amb_surf`) are run here on the seed-0 fixtures and their outputs are
committed as tests/golden/python_sibling.npz.  Two input precisions:

  * complex128 inputs (the f32 file samples widened exactly as caf_rust/src/utils.rs:19-32 does): caf.py then works in
    double precision end to end, so these vectors pin the fp64 surface magnitudes, apply_freq_shift values and
    Xcor magnitudes of the oracle and of the CUDA path at ~1e-13;
  * complex64 inputs, exactly as caf.py's own __main__ loads them (caf.py:129-130): a loose pin (~1e-5) on what the
    Python program itself prints.

Conventions (caf.py:12-13,145): row f is |correlate(shifted_needle, haystack, 'same')|, column j holds lag L/2 - j of
the Rust surface (caf_rust/src/caf/mod.rs:139-147), magnitude (not squared); peak delay = L//2 - argmax.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/caf_python")
import caf  # noqa: E402  (the reference module, unmodified)

DATA = os.path.join(HERE, "data")
FS = 48e3


def load(name, n=None):
    x = np.fromfile(os.path.join(DATA, name), dtype=np.complex64)
    return x if n is None else x[:n]


def main():
    out = {}
    # case A: the pair caf.py's __main__ uses (chirp_4), 6 doppler rows around the answer, complex128 inputs
    n64, h64 = load("chirp_4_raw.c64"), load("chirp_4_T+70samp_F+82.89Hz.c64", 4096)
    fa = np.array([80.0, 82.5, 82.9, 83.0, 83.5, 99.5])
    out["a_freqs"] = fa
    out["a_surf_c128"] = caf.amb_surf(n64.astype(np.complex128), h64.astype(np.complex128), fa, FS)
    # case B: the same call with complex64 inputs, as the Python program runs it (4 rows)
    fb = np.array([82.5, 83.0, 83.5, -20.0])
    out["b_freqs"] = fb
    out["b_surf_c64"] = caf.amb_surf(n64, h64, fb, FS)
    # case C: a short, ragged length (L = 1000) on another pair, complex128, negative and fractional shifts
    n7, h7 = load("chirp_7_raw.c64", 1000).astype(np.complex128), load("chirp_7_T+84samp_F+68.26Hz.c64", 1000).astype(np.complex128)
    fc = np.array([-92.75, 0.0, 68.25, 68.26, 70.0])
    out["c_freqs"] = fc
    out["c_surf_c128"] = caf.amb_surf(n7, h7, fc, FS)
    # case D: an ODD length (L = 999): pins the centring of scipy's 'same' window (column j <-> lag L//2 - j)
    fd = np.array([68.25, -5.0])
    out["d_freqs"] = fd
    out["d_surf_c128"] = caf.amb_surf(n7[:999], h7[:999], fd, FS)
    # apply_fdoa values (caf.py:28-33) == apply_freq_shift (mod.rs:46-65)
    x = n64[:1024].astype(np.complex128)
    out["shift_in"] = x
    out["shift_freq"] = np.array([77.77, -12.5])
    out["shift_out"] = np.stack([caf.apply_fdoa(x, f, FS) for f in out["shift_freq"]])
    # one plain correlation (caf.py:15-18): |correlate(a, b, 'same')|
    out["xcor_a"] = n7
    out["xcor_b"] = h7
    out["xcor_abs_same"] = caf.xcor(n7, h7)
    # what caf.py's __main__ reports for its benchmark grid (caf.py:134,144-146): tau_max, freq_max
    grid = np.arange(-100, 100, 0.5)
    surf = caf.amb_surf(n64, h64, grid[360:372], FS)            # rows 80.0 .. 85.5 Hz contain the maximum
    fmax, tmax = np.unravel_index(surf.argmax(), surf.shape)
    out["main_report"] = np.array([len(n64) // 2 - tmax, grid[360 + fmax]])
    # case E (round 2): ONE FULL 400-row surface of the README pair (chirp_0, caf_bench.rs' grid -100 .. 99.5 Hz), complex128
    # inputs: 400 x 4096 doubles = 13 MB, so what is committed is its SHA-256, every row's maximum and arg-maximum, and
    # 1 % of its cells (16 384 cells at positions drawn by a seeded generator), not the array
    import hashlib
    n0, h0 = load("chirp_0_raw.c64").astype(np.complex128), load("chirp_0_T+202samp_F+69.25Hz.c64", 4096).astype(np.complex128)
    grid = np.arange(-100000, 100000, 500) / 1e3                # caf_bench.rs:32-35
    full = caf.amb_surf(n0, h0, grid, FS)
    assert full.shape == (400, 4096) and full.dtype == np.float64
    rs = np.random.RandomState(20261018)
    flat = np.sort(rs.choice(full.size, size=16384, replace=False))
    out["e_freqs"] = grid
    out["e_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(full).tobytes()).digest(), dtype=np.uint8)
    out["e_cells_flat_index"] = flat.astype(np.int64)
    out["e_cells_value"] = full.ravel()[flat]
    out["e_row_max"] = full.max(axis=1)
    out["e_row_argmax"] = full.argmax(axis=1).astype(np.int64)
    out["e_surface_max"] = np.array([full.max()])
    path = os.path.join(HERE, "python_sibling.npz")
    np.savez(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "main_report", out["main_report"])


if __name__ == "__main__":
    main()
