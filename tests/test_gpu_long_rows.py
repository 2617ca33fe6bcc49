"""GPU parity for rows longer than 8192 delay cells (BASELINE config 3: 65 536 cells) — the four-step path.

Checker: the C oracle for power-of-two lengths, its numpy/pocketfft twin for the others (the C oracle's fallback for
non-power-of-two transforms is an O(n^2) long-double DFT, too slow at these sizes)."""
import numpy as np
import pytest

import caf_cookoff_b200 as caf
from caf_cookoff_b200 import api, generate as G
from conftest import FS, rel_max
from oracle import np_oracle as NO
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _pair(l, seed=0):
    p = G.pair(0, seed=seed, chirp_length=l)
    return G.as_inputs(p)


@pytest.mark.parametrize("l", [4097, 5000, 8192, 10000, 16384, 20001, 32768, 65536, 65537, 100000, 131072, 262144])
def test_long_rows_match_oracle(l):
    needle, hay = _pair(l)
    shifts = np.array([-50.0, 0.0, 12.5, 68.0, 69.25, 70.5])
    surf, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS)
    if l & (l - 1) == 0:
        osurf, opidx, opval = O.caf_surface(needle, hay, shifts, FS)
    else:
        osurf, opidx, opval = NO.caf_surface(needle, hay, shifts, FS, direct_phasor=True)
    assert surf.shape == (6, 2 * l)
    assert rel_max(surf, osurf) <= 1e-9
    assert np.array_equal(pidx.astype(np.int64), np.asarray(opidx).astype(np.int64))
    assert np.array_equal(pidx.astype(np.int64), np.argmax(surf, axis=1))
    assert (pk.freq_hz, int(pk.delay_idx)) == NO.find_peak(shifts, opidx, opval)


def test_config3_miniature_and_peak_only():
    """4096 doppler x 65536 delay is config 3; here 96 rows (two L2 chunks) of the same row length."""
    needle, hay = _pair(32768)
    shifts = np.linspace(-100.0, 100.0, 96, endpoint=False)
    surf, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS)
    osurf, opidx, opval = O.caf_surface(needle, hay, shifts, FS)
    assert rel_max(surf, osurf) <= 1e-9
    assert np.array_equal(pidx, opidx)
    assert (pk.freq_hz, int(pk.delay_idx)) == O.find_peak(shifts, opidx, opval)
    # peak-only entry (surface never materialised) agrees
    _, pidx2, pval2, pk2 = caf.surface_arrays(needle, hay, shifts, FS, want_surface=False)
    assert np.array_equal(pidx2, pidx) and np.array_equal(pval2, pval)
    assert (pk2.freq_hz, pk2.delay_idx, pk2.doppler_idx) == (pk.freq_hz, pk.delay_idx, pk.doppler_idx)


def test_long_rows_fp32():
    needle, hay = _pair(16384)
    shifts = np.array([0.0, 35.0, 69.25])
    surf, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS, variant=api._Variant32)
    osurf, opidx, _ = O.caf_surface(needle, hay, shifts, FS)
    assert surf.dtype == np.float32
    assert rel_max(surf.astype(np.float64), osurf) <= 1e-4
    assert int(pidx[2]) == int(opidx[2])


def test_two_level_rows_fp32():
    """complex64 through the fused two-level kernels (N = 2^18: top radix 2, middle radix 16)."""
    needle, hay = _pair(100000)
    shifts = np.array([0.0, 69.25])
    surf, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS, variant=api._Variant32)
    osurf, opidx, _ = NO.caf_surface(needle, hay, shifts, FS, direct_phasor=True)
    assert surf.dtype == np.float32 and surf.shape == (2, 200000)
    assert rel_max(surf.astype(np.float64), osurf) <= 1e-4
    assert int(pidx[1]) == int(opidx[1])


def test_long_row_delay_and_doppler_properties():
    """The generator puts s1 `lag` samples and `foffset` Hz away from s0: the peak must land there."""
    p = G.pair(0, seed=3, chirp_length=65536)
    needle, hay = G.as_inputs(p)
    step = 0.5
    f0 = round(p.foffset_hz / step) * step
    shifts = np.array([f0 - step, f0, f0 + step])
    _, _, _, pk = caf.surface_arrays(needle, hay, shifts, FS, want_surface=False)
    assert int(pk.delay_idx) == p.lag
    assert pk.freq_hz == f0


def test_config5_row_length_peak_only():
    """BASELINE config 5 row length (2^20 delay cells), peak only: the surface is never materialised."""
    p = G.pair(0, seed=1, chirp_length=1 << 19)
    needle, hay = G.as_inputs(p)
    step = 0.25
    f0 = round(p.foffset_hz / step) * step
    shifts = np.array([f0 - step, f0, f0 + step, f0 + 2 * step])
    _, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS, want_surface=False)
    _, opidx, opval = O.caf_surface(needle, hay, shifts, FS, want_surface=False)
    assert np.array_equal(pidx, opidx)
    assert rel_max(pval, opval) <= 1e-9
    assert (pk.freq_hz, int(pk.delay_idx)) == O.find_peak(shifts, opidx, opval)
    assert abs(int(pk.delay_idx) - p.lag) <= 16     # the chirp's delay-doppler ridge moves the peak a few samples


def test_too_long_is_unsupported():
    z = np.zeros((1 << 19) + 1, dtype=complex)
    with pytest.raises(caf.CafError) as e:
        caf.surface_arrays(z, z, [0.0], FS)
    assert e.value.status == -3


@pytest.mark.parametrize("l", [8192, 131072])
def test_long_rows_beyond_nyquist_and_non_finite_doppler(l):
    """The long-row kernels have their own phasor (caf_large_spread_top / caf_large_spread2): shifts beyond fs/2 alias,
    a NaN or infinite shift poisons its row and never wins (mod.rs:55-60, 148-151) — one- and two-level rows."""
    needle, hay = _pair(l)
    shifts = np.array([69.25, 69.25 + FS, np.nan, 69.25 - 2 * FS, np.inf, FS / 2.0])
    surf, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS)
    osurf, opidx, opval = O.caf_surface(needle, hay, shifts, FS)
    for r in (2, 4):
        assert np.isnan(osurf[r]).all() and np.isnan(surf[r]).all()
        assert (int(pidx[r]), float(pval[r])) == (0, 0.0) == (int(opidx[r]), float(opval[r]))
    fin = [0, 1, 3, 5]
    assert rel_max(surf[fin], osurf[fin]) <= 1e-9
    pk_rows = [0, 1, 3]          # the fs/2 row holds rounding noise only (1e-9 of the maximum): its argmax is not defined
    assert np.array_equal(pidx[pk_rows].astype(np.int64), np.asarray(opidx)[pk_rows].astype(np.int64))
    assert rel_max(surf[1], surf[0]) <= 1e-9 and rel_max(surf[3], surf[0]) <= 1e-9
    assert int(pk.delay_idx) == int(opidx[0]) and pk.freq_hz in (69.25, 69.25 + FS, 69.25 - 2 * FS)
    _, _, _, pk2 = caf.surface_arrays(needle, hay, shifts, FS, want_surface=False)
    assert (pk2.freq_hz, int(pk2.delay_idx), int(pk2.doppler_idx)) == (pk.freq_hz, int(pk.delay_idx), int(pk.doppler_idx))
