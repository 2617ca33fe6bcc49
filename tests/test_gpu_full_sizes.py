"""BASELINE configs 3, 4 and 5 at their FULL sizes, checked through properties that do not need the oracle to finish
(it would take minutes to hours at these sizes): the generator's planted (lag, doppler), invariance of the answer under
sharding and under restriction of the grid, bit-exact scaling by a power of two, equality of repeated pairs."""
import numpy as np
import pytest

import caf_cookoff_b200 as caf
from caf_cookoff_b200 import api, generate as G
from conftest import FS

pytestmark = pytest.mark.gpu


def _resolve_halves(needle, hay, shifts, world):
    words = []
    for r in range(world):
        lo, hi = len(shifts) * r // world, len(shifts) * (r + 1) // world
        _, _, _, pk = caf.surface_arrays(needle, hay, shifts[lo:hi], FS, want_surface=False)
        words.append(api.peak_pack(pk, lo))
    return api.peak_resolve(np.stack(words))


def test_config3_full_size_properties():
    """4096 doppler x 65536 delay, peak and row peaks (the 2.1 GB surface stays on the GPU)."""
    p = G.pair(0, seed=0, chirp_length=32768)
    needle, hay = G.as_inputs(p)
    shifts = np.linspace(-100.0, 100.0, 4096, endpoint=False)
    _, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS, want_surface=False)
    assert pidx.shape == (4096,)
    # the planted offset: the peak row is the grid point next to the generator's doppler, the delay is its lag up to
    # the chirp's delay-doppler ridge
    assert abs(pk.freq_hz - p.foffset_hz) <= 200.0 / 4096
    assert abs(int(pk.delay_idx) - p.lag) <= 4
    assert int(pk.doppler_idx) == int(np.argmax(pval)) and int(pk.delay_idx) == int(pidx[int(pk.doppler_idx)])
    # rows sharded over 8 "ranks" + packed maxloc == the unsharded answer, bit for bit
    out = _resolve_halves(needle, hay, shifts, 8)
    assert (out.value, out.freq_hz, out.doppler_idx, out.delay_idx) == (pk.value, pk.freq_hz, pk.doppler_idx, pk.delay_idx)
    # |xcor|^2 scales with |c|^2 of the haystack: exact for c = 2
    _, pidx2, pval2, _ = caf.surface_arrays(needle, 2.0 * hay, shifts, FS, want_surface=False)
    assert np.array_equal(pidx2, pidx) and np.array_equal(pval2, 4.0 * pval)


def test_config4_full_size_properties():
    """4096 independent pairs x (400 x 8192), peaks only: every copy of a pair gives that pair's known answer."""
    import os
    from conftest import DATA
    names = sorted(os.listdir(DATA))
    ns = np.stack([caf.read_file_c64(os.path.join(DATA, f"chirp_{i}_raw.c64")) for i in range(10)])
    hs = np.stack([caf.read_file_c64(os.path.join(DATA, [n for n in names if n.startswith(f"chirp_{i}_T")][0]))[:4096] for i in range(10)])
    idx = np.arange(4096) % 10
    shifts = caf.bench_shifts()
    _, pidx, pval, peaks = caf.batch_arrays(ns[idx], hs[idx], shifts, FS, want_surface=False)
    assert pidx.shape == (4096, 400)
    single = [caf.surface_arrays(ns[i], hs[i], shifts, FS, want_surface=False) for i in range(10)]
    for j in range(4096):
        _, spi, spv, spk = single[j % 10]
        assert (peaks[j].freq_hz, peaks[j].delay_idx, peaks[j].doppler_idx, peaks[j].value) == \
               (spk.freq_hz, spk.delay_idx, spk.doppler_idx, spk.value)
    assert np.array_equal(pidx, np.stack([single[j % 10][1] for j in range(4096)]))
    assert np.array_equal(pval, np.stack([single[j % 10][2] for j in range(4096)]))
    assert (peaks[0].freq_hz, int(peaks[0].delay_idx)) == (69.0, 202)          # chirp_0 on the bench grid


def test_config5_full_size_properties():
    """16384 doppler x 2^20 delay, peak only (3.7 TFLOP of algorithmic work; the surface would be 137 GB)."""
    p = G.pair(0, seed=1, chirp_length=1 << 19)
    needle, hay = G.as_inputs(p)
    shifts = np.linspace(-100.0, 100.0, 16384, endpoint=False)
    _, pidx, pval, pk = caf.surface_arrays(needle, hay, shifts, FS, want_surface=False)
    row = int(pk.doppler_idx)
    assert abs(pk.freq_hz - p.foffset_hz) <= 200.0 / 16384 * 2
    assert abs(int(pk.delay_idx) - p.lag) <= 16
    assert row == int(np.argmax(pval))
    # restricting the grid to a window around the winner does not change the winner or its row values (rows are independent)
    lo, hi = max(row - 37, 0), min(row + 40, 16384)
    _, pidx_w, pval_w, pk_w = caf.surface_arrays(needle, hay, shifts[lo:hi], FS, want_surface=False)
    assert np.array_equal(pidx_w, pidx[lo:hi]) and np.array_equal(pval_w, pval[lo:hi])
    assert (pk_w.freq_hz, pk_w.delay_idx, pk_w.value, int(pk_w.doppler_idx) + lo) == (pk.freq_hz, pk.delay_idx, pk.value, row)
    # two-rank sharding resolves to the same answer
    out = _resolve_halves(needle, hay, shifts[lo:hi], 2)
    assert (out.value, out.freq_hz, out.delay_idx, int(out.doppler_idx) + lo) == (pk.value, pk.freq_hz, pk.delay_idx, row)
