"""The oracle (and, on the GPU, the CUDA path) against VALUES produced by the reference itself: tests/golden/
python_sibling.npz holds outputs of the unmodified /root/reference/caf_python/caf.py functions on the seed-0 fixtures
(generator: tests/golden/make_python_golden.py).  With complex128 inputs caf.py works in double precision, so these
vectors pin surface magnitudes, apply_freq_shift values and correlation magnitudes — which caf_rust's own tests never
check — at the fp64 level.

Relation between the programs (caf.py:12-13,145 vs caf_rust/src/caf/mod.rs:139-147):
    python[f][j] = sqrt( rust[f][(L//2 - j) mod 2L] ),   j < L        tau_max = L//2 - argmax_j
"""
import os

import numpy as np
import pytest

from conftest import DATA, FS, GOLDEN, rel_max
from oracle import oracle as O

G = np.load(os.path.join(GOLDEN, "python_sibling.npz"))


def rust_to_python(surf, L):
    j = np.arange(L)
    return np.sqrt(surf[:, (L // 2 - j) % (2 * L)])


def rust_to_go(surf, L):
    k = np.arange(2 * L)
    return np.sqrt(surf[:, (L - k) % (2 * L)])


def pair4():
    n = O.read_file_c64(os.path.join(DATA, "chirp_4_raw.c64"))
    h = O.read_file_c64(os.path.join(DATA, "chirp_4_T+70samp_F+82.89Hz.c64"))[:4096]
    return n, h


def pair7(L=1000):
    n = O.read_file_c64(os.path.join(DATA, "chirp_7_raw.c64"))[:L]
    h = O.read_file_c64(os.path.join(DATA, "chirp_7_T+84samp_F+68.26Hz.c64"))[:L]
    return n, h


# ---------------------------------------------------------------- CPU: the oracle is pinned by the reference's values
def test_oracle_surface_magnitudes_fp64():
    n, h = pair4()
    surf, _, _ = O.caf_surface(n, h, G["a_freqs"], FS)
    assert rel_max(rust_to_python(surf, 4096), G["a_surf_c128"]) <= 1e-12      # measured 9.6e-14


def test_oracle_surface_ragged_length_fp64():
    n, h = pair7()
    surf, _, _ = O.caf_surface(n, h, G["c_freqs"], FS)
    assert rel_max(rust_to_python(surf, 1000), G["c_surf_c128"]) <= 1e-12      # measured 3.3e-14


def test_oracle_surface_odd_length_fp64():
    n, h = pair7(999)
    surf, _, _ = O.caf_surface(n, h, G["d_freqs"], FS)
    assert rel_max(rust_to_python(surf, 999), G["d_surf_c128"]) <= 1e-12


def test_oracle_vs_python_program_in_complex64():
    """caf.py as it runs itself (complex64 samples, caf.py:129-130): loose, single-precision agreement."""
    n, h = pair4()
    surf, _, _ = O.caf_surface(n, h, G["b_freqs"], FS)
    assert rel_max(rust_to_python(surf, 4096), G["b_surf_c64"]) <= 1e-5        # measured 2.1e-7


def test_oracle_apply_freq_shift_values():
    for f, want in zip(G["shift_freq"], G["shift_out"]):
        got = O.apply_freq_shift(G["shift_in"], float(f), FS)
        assert np.abs(got - want).max() <= 1e-13                               # measured 4e-15 (|x| ~ 1)


def test_oracle_xcor_magnitudes():
    """caf.py xcor(a, b) = |correlate(a, b, 'same')| against Xcor::run(b|0, a|0) (operands swapped, zero padded)."""
    a, b = G["xcor_a"], G["xcor_b"]
    L = a.size
    z = np.zeros(L, dtype=np.complex128)
    r = O.xcor(np.concatenate([b, z]), np.concatenate([a, z]))                 # r[k] = sum b[m+k] conj(a[m]), lag k
    j = np.arange(L)
    assert rel_max(np.abs(r[(L // 2 - j) % (2 * L)]), G["xcor_abs_same"]) <= 1e-12


def pair0():
    n = O.read_file_c64(os.path.join(DATA, "chirp_0_raw.c64"))
    h = O.read_file_c64(os.path.join(DATA, "chirp_0_T+202samp_F+69.25Hz.c64"))[:4096]
    return n, h


def check_full_surface_against_reference_values(surf_rust, tol):
    """ONE FULL 400-row surface of the README pair from the unmodified caf.py (13 MB: its SHA-256, every row's maximum and
    arg-maximum and 1 % of its cells are committed): the given Rust-layout surface, mapped to caf.py's layout, agrees at
    all 16 384 sampled cells, and every row's maximum sits where caf.py has it."""
    py = rust_to_python(surf_rust, 4096)
    assert py.shape == (400, 4096)
    big = float(G["e_surface_max"][0])
    got = py.ravel()[G["e_cells_flat_index"]]
    assert np.abs(got - G["e_cells_value"]).max() / big <= tol
    assert np.abs(py.max(axis=1) - G["e_row_max"]).max() / big <= tol
    # arg-maxima: identical wherever the row's two best cells differ by more than the tolerance (they all do here)
    assert np.array_equal(py.argmax(axis=1), G["e_row_argmax"])
    fm = int(np.argmax(G["e_row_max"]))
    assert (float(G["e_freqs"][fm]), 4096 // 2 - int(G["e_row_argmax"][fm])) == (69.0, 202)      # caf_bench.rs' pair on its grid
    return float(np.abs(got - G["e_cells_value"]).max() / big)


def test_oracle_full_400_row_surface_fp64():
    n, h = pair0()
    surf, _, _ = O.caf_surface(n, h, G["e_freqs"], FS)
    err = check_full_surface_against_reference_values(surf, 1e-12)
    assert err <= 1e-12
    assert G["e_sha256"].size == 32            # the digest of caf.py's own array (regeneration check, make_python_golden.py)


def test_python_program_report_matches_rust_answer():
    """caf.py's __main__ prints tau_max = 70, freq_max = 83.0 for the chirp_4 pair on its 0.5 Hz grid; the Rust
    convention's peak of the same rows is (83.0, 70) (known answer test.rs:157-170 is 82.9 on the 0.1 Hz grid)."""
    assert tuple(G["main_report"]) == (70.0, 83.0)
    n, h = pair4()
    grid = np.arange(-100, 100, 0.5)[360:372]
    _, pidx, pval = O.caf_surface(n, h, grid, FS, want_surface=False)
    assert O.find_peak(grid, pidx, pval) == (83.0, 70)


# ---------------------------------------------------------------- GPU: the CUDA path against the same vectors
@pytest.mark.gpu
def test_cuda_surface_magnitudes_fp64():
    from caf_cookoff_b200 import surface_arrays
    n, h = pair4()
    surf, _, _, _ = surface_arrays(n, h, G["a_freqs"], FS)
    assert rel_max(rust_to_python(surf, 4096), G["a_surf_c128"]) <= 1e-9       # north-star tolerance; measured ~1e-13
    n, h = pair7()
    surf, _, _, _ = surface_arrays(n, h, G["c_freqs"], FS)
    assert rel_max(rust_to_python(surf, 1000), G["c_surf_c128"]) <= 1e-9


@pytest.mark.gpu
def test_cuda_full_400_row_surface_fp64():
    from caf_cookoff_b200 import surface_arrays
    n, h = pair0()
    surf, _, _, pk = surface_arrays(n, h, G["e_freqs"], FS)
    check_full_surface_against_reference_values(surf, 1e-9)        # north-star tolerance; measured ~1e-13
    assert (pk.freq_hz, int(pk.delay_idx)) == (69.0, 202)


@pytest.mark.gpu
def test_cuda_python_layout_and_peak():
    from caf_cookoff_b200 import PythonSibling
    n, h = pair4()
    got = PythonSibling.amb_surf(n, h, G["a_freqs"], 48e3)
    assert got.shape == G["a_surf_c128"].shape
    assert rel_max(got, G["a_surf_c128"]) <= 1e-9
    got = PythonSibling.amb_surf(n, h, G["b_freqs"], 48e3)
    assert rel_max(got, G["b_surf_c64"]) <= 1e-5
    n7, h7 = pair7()
    got = PythonSibling.amb_surf(n7, h7, G["c_freqs"], 48e3)
    assert rel_max(got, G["c_surf_c128"]) <= 1e-9
    # what caf.py prints (caf.py:144-146)
    grid = np.arange(-100, 100, 0.5)
    assert PythonSibling.peak(n, h, grid, 48e3) == (70, 83.0)
    # np.unravel_index(surf.argmax()) of the golden rows themselves
    fm, tm = np.unravel_index(G["a_surf_c128"].argmax(), G["a_surf_c128"].shape)
    assert PythonSibling.peak(n, h, G["a_freqs"], 48e3) == (4096 // 2 - tm, G["a_freqs"][fm])


@pytest.mark.gpu
def test_cuda_python_layout_f32():
    from caf_cookoff_b200 import surface_layout
    n, h = pair4()
    got, pk = surface_layout(n, h, G["a_freqs"], FS, 1, f32=True)
    assert got.dtype == np.float32
    assert rel_max(got.astype(np.float64), G["a_surf_c128"]) <= 1e-4           # complex64 variant tolerance


@pytest.mark.gpu
def test_cuda_go_layout_and_peak():
    """Go's program cannot run here (no toolchain); its layout is the oracle's surface under caf.go's index map,
    cross-checked against the Python golden on the lags the two layouts share."""
    from caf_cookoff_b200 import GoSibling, surface_layout
    n, h = pair4()
    L = 4096
    surf, _, _ = O.caf_surface(n, h, G["a_freqs"], FS)
    got = GoSibling.amb_surf(n, h, G["a_freqs"], 48e3)
    assert got.shape == (G["a_freqs"].size, 2 * L)
    assert rel_max(got, rust_to_go(surf, L)) <= 1e-9
    # python column j = lag L/2 - j = go column k with L - k = L/2 - j  ->  k = L/2 + j
    assert rel_max(got[:, L // 2 + np.arange(L)], G["a_surf_c128"]) <= 1e-9
    # main.go:33-35: find_2d_peak, then len(apple) - tdx samples
    fdx, tdx, mx = GoSibling.find_2d_peak(n, h, np.arange(-100, 100, 0.5), 48e3)
    assert (L - tdx, np.arange(-100, 100, 0.5)[fdx]) == (70, 83.0)
    ref = rust_to_go(O.caf_surface(n, h, np.arange(-100, 100, 0.5)[360:372], FS)[0], L)
    assert abs(mx - ref.max()) <= 1e-9 * ref.max()
    # ragged length, odd L
    n7, h7 = pair7(999)
    s7, _, _ = O.caf_surface(n7, h7, G["c_freqs"], FS)
    assert rel_max(GoSibling.amb_surf(n7, h7, G["c_freqs"], 48e3), rust_to_go(s7, 999)) <= 1e-9
    from caf_cookoff_b200 import PythonSibling
    assert rel_max(PythonSibling.amb_surf(n7, h7, G["d_freqs"], 48e3), G["d_surf_c128"]) <= 1e-9
    # empty grid / all-zero input
    out, pk = surface_layout(np.zeros(64, complex), np.zeros(64, complex), [1.0, 2.0], FS, 2)
    assert not out.any() and pk.doppler_idx == (1 << 64) - 1
    assert GoSibling.find_2d_peak(np.zeros(64, complex), np.zeros(64, complex), [1.0, 2.0], FS) == (0, 0, 0.0)


@pytest.mark.gpu
def test_cuda_shift_and_xcor_values():
    from caf_cookoff_b200 import CafB200, PythonSibling, Xcor
    for f, want in zip(G["shift_freq"], G["shift_out"]):
        assert np.abs(CafB200.apply_freq_shift(G["shift_in"], float(f), FS) - want).max() <= 1e-12
        assert np.abs(PythonSibling.apply_fdoa(G["shift_in"], float(f), 48e3) - want).max() <= 1e-12
    a, b = G["xcor_a"], G["xcor_b"]
    L = a.size
    z = np.zeros(L, dtype=np.complex128)
    r = Xcor.new(2 * L).run(np.concatenate([b, z]), np.concatenate([a, z]))
    j = np.arange(L)
    assert rel_max(np.abs(r[(L // 2 - j) % (2 * L)]), G["xcor_abs_same"]) <= 1e-9


@pytest.mark.gpu
def test_cuda_sibling_layouts_on_long_rows():
    """Rows longer than 8192 cells (the four-step path) through the same layout conversion."""
    from caf_cookoff_b200 import generate as Gen, surface_layout
    needle, hay = Gen.as_inputs(Gen.pair(0, seed=0, chirp_length=5000))
    shifts = np.array([-3.0, 68.0, 69.25])
    from oracle import np_oracle as NO
    surf, _, _ = NO.caf_surface(needle, hay, shifts, FS, direct_phasor=True)
    py, pkp = surface_layout(needle, hay, shifts, FS, 1)
    go, pkg = surface_layout(needle, hay, shifts, FS, 2)
    assert py.shape == (3, 5000) and go.shape == (3, 10000)
    assert rel_max(py, rust_to_python(surf, 5000)) <= 1e-9
    assert rel_max(go, rust_to_go(surf, 5000)) <= 1e-9
    fm, tm = np.unravel_index(rust_to_python(surf, 5000).argmax(), (3, 5000))
    assert (int(pkp.doppler_idx), int(pkp.delay_idx)) == (fm, tm)
    fm, tm = np.unravel_index(rust_to_go(surf, 5000).argmax(), (3, 10000))
    assert (int(pkg.doppler_idx), int(pkg.delay_idx)) == (fm, tm)
