"""CPU tests of the oracle (oracle/): pinned on the reference's own known answers and cross-checked three ways.

These are the parity pins of the whole repo: the reference crate cannot be built here (no cargo, no libfftw3), so the
C restatement is validated against (1) every (freq, samp_idx) assertion of caf_rust/tests/test.rs on the reference's
seed-0 fixtures, (2) a numpy/pocketfft twin, (3) a long-double direct evaluation without any FFT.
"""
import os

import numpy as np
import pytest

from conftest import DATA, FS, load_case, rel_max
from oracle import np_oracle as NO
from oracle import oracle as O


def test_known_answers_every_chirp(known_answers):
    """caf_rust/tests/test.rs:15-317 — all ten (freq, samp_idx) pairs, exact equality like assert_eq!."""
    for case in known_answers:
        needle, hay, shifts = load_case(case)
        _, pidx, pval = O.caf_surface(needle, hay, shifts, FS, want_surface=False)
        freq, idx = O.find_peak(shifts, pidx, pval)
        assert freq == case["freq"], case
        assert idx == case["samp_idx"], case


def test_threadpool_variant_matches_serial(chirp0):
    """CafRustFFTThreadpool (mod.rs:391-461) computes the same rows as CafRustFFT (mod.rs:121-166)."""
    needle, hay = chirp0
    shifts = O.gen_float_shifts(60.0, 80.0, 0.25)
    s1, i1, v1 = O.caf_surface(needle, hay, shifts, FS)
    s2, i2, v2 = O.caf_surface(needle, hay, shifts, FS, threads=4)
    assert np.array_equal(s1, s2) and np.array_equal(i1, i2) and np.array_equal(v1, v2)


def test_bench_grid_peak(chirp0):
    """The benchmark grid (caf_bench.rs:30-35): 400 rows, peak at row 338 = 69.0 Hz, delay 202."""
    needle, hay = chirp0
    shifts = O.bench_shifts()
    assert shifts.size == 400 and shifts[0] == -100.0 and shifts[-1] == 99.5
    _, pidx, pval = O.caf_surface(needle, hay, shifts, FS, want_surface=False)
    assert O.find_peak(shifts, pidx, pval) == (69.0, 202)
    assert int(np.argmax(pval)) == 338
    assert abs(pval.max() - 902.604145) < 1e-5


def test_c_oracle_vs_numpy_twin(chirp0):
    needle, hay = chirp0
    shifts = O.gen_float_shifts(65.0, 73.0, 0.5)
    s_c, i_c, v_c = O.caf_surface(needle, hay, shifts, FS)
    s_n, i_n, v_n = NO.caf_surface(needle, hay, shifts, FS, direct_phasor=False)   # same recursion as the reference
    assert rel_max(s_c, s_n) < 1e-14
    assert np.array_equal(i_c, i_n)
    s_d, i_d, _ = NO.caf_surface(needle, hay, shifts, FS, direct_phasor=True)      # closed-form phasor
    assert rel_max(s_c, s_d) < 1e-12      # the reference's acc *= shift recursion drifts ~2e-13
    assert np.array_equal(i_c, i_d)


def test_direct_long_double_cells(chirp0):
    """No FFT, no phasor recursion: out[k] = sum_m hay[m+k] conj(needle[m] e^{j 2 pi f m/fs}) in long double."""
    needle, hay = chirp0
    f = 69.0
    surf, _, _ = O.caf_surface(needle, hay, [f], FS)
    lags = np.array([0, 1, 201, 202, 203, 4095, 4096, 4097, 8000, 8191], dtype=np.uint64)
    truth = O.direct_cells(needle, hay, f, FS, lags)
    assert np.abs(truth - surf[0][lags]).max() / surf.max() < 1e-12


def test_apply_freq_shift_recursion_and_closed_form():
    rng = np.random.default_rng(7)
    x = rng.normal(size=4096) + 1j * rng.normal(size=4096)
    y_c = O.apply_freq_shift(x, 77.77, FS)                      # caf_bench.rs:171-179 uses 77.77 Hz
    y_np = NO.apply_freq_shift(x, 77.77, FS)
    assert rel_max(y_c, y_np) < 1e-15
    assert rel_max(y_c, NO.apply_freq_shift_direct(x, 77.77, FS)) < 1e-12
    assert np.array_equal(O.apply_freq_shift(x, 0.0, FS), x)     # zero shift is the identity
    assert O.apply_freq_shift(np.zeros(0, dtype=complex), 5.0, FS).size == 0


@pytest.mark.parametrize("n", [1, 2, 8, 37, 100, 1024, 8192])
def test_xcor_matches_numpy(n):
    """Xcor::run (xcor_rustfft.rs:51-78) for power-of-two and other lengths."""
    rng = np.random.default_rng(n)
    a = rng.normal(size=n) + 1j * rng.normal(size=n)
    b = rng.normal(size=n) + 1j * rng.normal(size=n)
    assert rel_max(O.xcor(a, b), NO.xcor(a, b)) < 1e-13


def test_xcor_is_circular_correlation():
    rng = np.random.default_rng(3)
    n = 64
    a = rng.normal(size=n) + 1j * rng.normal(size=n)
    b = rng.normal(size=n) + 1j * rng.normal(size=n)
    direct = np.array([sum(a[(m + k) % n] * np.conj(b[m]) for m in range(n)) for k in range(n)])
    assert rel_max(O.xcor(a, b), direct) < 1e-13


def test_length_mismatch_panics():
    """xcor_rustfft.rs:54-55 assert!(a.len() == self.n)."""
    with pytest.raises(AssertionError):
        O.xcor(np.zeros(8, complex), np.zeros(9, complex))
    with pytest.raises(AssertionError):
        O.caf_surface(np.zeros(8, complex), np.zeros(9, complex), [0.0], FS)


def test_find_peak_semantics():
    """mod.rs:31-42: strict > from a dummy row (0.0, peak 0.0): first maximal row wins, empty/all-zero -> (0.0, 0)."""
    f = np.array([1.0, 2.0, 3.0, 4.0])
    assert O.find_peak(f, [5, 6, 7, 8], [1.0, 9.0, 9.0, 2.0]) == (2.0, 6)     # tie -> first row
    assert O.find_peak(f, [5, 6, 7, 8], [0.0, 0.0, 0.0, 0.0]) == (0.0, 0)     # nothing beats the dummy row
    assert O.find_peak(f[:0], [], []) == (0.0, 0)
    assert O.find_peak(f, [5, 6, 7, 8], [np.nan, 1.0, np.nan, 0.5]) == (2.0, 6)  # NaN never wins a strict >


def test_row_argmax_first_maximum_and_zero_rows():
    """mod.rs:141-153: running max from 0.0 with strict > -> first maximal index; an all-zero row reports (0, 0.0)."""
    z = np.zeros(16, dtype=complex)
    _, pidx, pval = O.caf_surface(z, z, [0.0, 1.0], FS)
    assert list(pidx) == [0, 0] and list(pval) == [0.0, 0.0]
    # a delta needle against a haystack with two equal taps: lags 3 and 9 tie exactly, the first must win
    needle = np.zeros(16, dtype=complex); needle[0] = 1.0
    hay = np.zeros(16, dtype=complex); hay[3] = 2.0; hay[9] = 2.0
    surf, pidx, pval = O.caf_surface(needle, hay, [0.0], FS)
    assert surf[0][3] == surf[0][9] == pval[0]
    assert int(pidx[0]) == 3


def test_ragged_lengths_against_numpy_twin():
    """The reference accepts any length (RustFFT plans any n); the oracle falls back to a long-double DFT."""
    rng = np.random.default_rng(11)
    for l in (1, 3, 17, 100):
        needle = rng.normal(size=l) + 1j * rng.normal(size=l)
        hay = rng.normal(size=l) + 1j * rng.normal(size=l)
        shifts = [-40.0, 0.0, 12.5]
        s_c, i_c, _ = O.caf_surface(needle, hay, shifts, FS)
        s_n, i_n, _ = NO.caf_surface(needle, hay, shifts, FS, direct_phasor=False)
        assert s_c.shape == (3, 2 * l)
        assert rel_max(s_c, s_n) < 1e-12
        assert np.array_equal(i_c, i_n)


def test_gen_float_shifts_semantics():
    """tests/test.rs:335-352: integer milli-Hz half-open range."""
    assert O.gen_float_shifts(-100.0, 100.0, 0.25).size == 800
    assert O.gen_float_shifts(30.0, 35.0, 0.05).size == 100
    g = O.gen_float_shifts(80.0, 100.0, 0.1)
    assert g.size == 200 and g[29] == 82.9
    assert O.gen_float_shifts(-50.0, 50.0, 1.0)[86] == 36.0


def test_read_file_c64_widening():
    """utils.rs:10-35: f32 LE pairs -> f64 pairs, exactly."""
    path = os.path.join(DATA, "chirp_0_raw.c64")
    raw = np.fromfile(path, dtype=np.complex64)
    x = O.read_file_c64(path)
    assert x.dtype == np.complex128 and x.size == 4096
    assert np.array_equal(x, raw.astype(np.complex128))


def test_non_finite_doppler_rows_never_win(chirp0):
    """from_polar(1, NaN) poisons the whole row (mod.rs:55-60); NaN never passes the strict > (mod.rs:148-151, 36-40)."""
    needle, hay = chirp0
    shifts = np.array([np.nan, 69.0, np.inf, -np.inf, 69.25])
    surf, pidx, pval = O.caf_surface(needle, hay, shifts, 48000)
    for r in (0, 2, 3):
        assert np.isnan(surf[r]).all() and (int(pidx[r]), float(pval[r])) == (0, 0.0)
    assert O.find_peak(shifts, pidx, pval) == (69.25, 202)
    assert O.find_peak(shifts[[0, 2]], pidx[[0, 2]], pval[[0, 2]]) == (0.0, 0)
