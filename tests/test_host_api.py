"""CPU tests of the host-side mirror of caf_rust's API (no GPU work): I/O formats, doppler grids, find_peak on rows,
argument checks that mirror the reference's panics."""
import os

import numpy as np
import pytest

import caf_cookoff_b200 as caf
from caf_cookoff_b200 import api
from conftest import DATA
from oracle import oracle as O


def test_read_file_c64_matches_reference_semantics(tmp_path):
    path = os.path.join(DATA, "chirp_3_raw.c64")
    x = caf.read_file_c64(path)
    assert x.dtype == np.complex128 and x.size == 4096
    assert np.array_equal(x, O.read_file_c64(path))
    bad = tmp_path / "bad.c64"
    bad.write_bytes(b"\0" * 12)
    with pytest.raises(ValueError):
        caf.read_file_c64(str(bad))


def test_write_file_binary_roundtrip(tmp_path):
    """utils.rs:39-63: numpy complex128 compatible."""
    x = (np.arange(10) + 1j * np.arange(10)[::-1]).astype(np.complex128)
    p = tmp_path / "out.bin"
    caf.write_file_binary(x, str(p))
    assert os.path.getsize(p) == 160
    assert np.array_equal(np.fromfile(p, dtype=np.complex128), x)


def test_shift_grids():
    assert np.array_equal(caf.gen_float_shifts(-100.0, 100.0, 0.25), O.gen_float_shifts(-100.0, 100.0, 0.25))
    assert np.array_equal(caf.bench_shifts(), O.bench_shifts())
    assert caf.gen_float_shifts(30.0, 35.0, 0.05)[43] == 32.15


def test_find_peak_on_rows():
    rows = [api.CafSurfaceRow(1.0, None, 5, 0.5), api.CafSurfaceRow(2.0, None, 6, 3.0),
            api.CafSurfaceRow(3.0, None, 7, 3.0), api.CafSurfaceRow(4.0, None, 8, float("nan"))]
    assert caf.CafB200.find_peak(rows) == (2.0, 6)
    assert caf.CafB200.find_peak([]) == (0.0, 0)
    assert caf.CafRustFFTThreadpool.find_peak(rows) == (2.0, 6)


def test_all_seven_strategy_names_exist():
    """caf_bench.rs:12-19 / tests/test.rs use these names."""
    for name in ["CafFFTW", "CafRustFFT", "CafRustFFTRayon", "CafRustFFTIter", "CafRustFFTIterRayon",
                 "CafRustFFTThreads", "CafRustFFTThreadpool"]:
        cls = getattr(caf, name)
        assert issubclass(cls, caf.CafSurface)
        for fn in ("caf_surface", "find_peak", "apply_freq_shift"):
            assert callable(getattr(cls, fn))


def test_length_mismatch_panics_before_any_gpu_work():
    with pytest.raises(caf.CafPanic):
        caf.CafB200.caf_surface(np.zeros(8, complex), np.zeros(9, complex), [0.0], 48000)
    with pytest.raises(caf.CafPanic):
        caf.Xcor.new(8).run(np.zeros(8, complex), np.zeros(7, complex))
    with pytest.raises(caf.CafPanic):
        caf.CafRustFFTIter.caf_surface(np.zeros(0, complex), np.zeros(0, complex), [1.0], 48000)  # xcor_mag[0]


def test_peak_pack_resolve_matches_find_peak_model():
    """caf_b200_peak_pack / _resolve (pure host helpers) against a direct model of find_peak (mod.rs:31-42) over the
    concatenated rows of all ranks: strict > from a dummy 0.0 row, so the first row holding the maximum wins, a rank
    whose shard found nothing contributes nothing, and an all-zero surface resolves to the dummy row."""
    from hypothesis import given, settings, strategies as st
    from caf_cookoff_b200 import _lib, api

    vals = st.sampled_from([0.0, 1.0, 2.5, 2.5, 7.0, 7.0, 1e-300, 3.0])

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.lists(st.tuples(vals, st.integers(0, 8191)), min_size=0, max_size=5), min_size=1, max_size=8))
    def check(shards):
        words, rows, off = [], [], 0
        for shard in shards:
            # this rank's own find_peak over its rows
            best, brow, bdelay = 0.0, None, 0
            for r, (v, dly) in enumerate(shard):
                rows.append((v, dly))
                if v > best:
                    best, brow, bdelay = v, r, dly
            pk = _lib.Peak()
            if brow is None:
                pk.value, pk.freq_hz, pk.doppler_idx, pk.delay_idx = 0.0, 0.0, api.UINT64_MAX, 0
            else:
                pk.value, pk.freq_hz, pk.doppler_idx, pk.delay_idx = best, float(off + brow), brow, bdelay
            words.append(api.peak_pack(pk, off))
            off += len(shard)
        got = api.peak_resolve(np.stack(words))
        best, brow = 0.0, None
        for r, (v, dly) in enumerate(rows):
            if v > best:
                best, brow = v, r
        if brow is None:
            assert got.doppler_idx == api.UINT64_MAX and got.value == 0.0 and got.delay_idx == 0
        else:
            assert (got.value, int(got.doppler_idx), int(got.delay_idx), got.freq_hz) == (best, brow, rows[brow][1], float(brow))

    check()


def test_sibling_adapters_refuse_a_sample_rate_the_u32_cannot_hold():
    """caf.py:28 / caf.go:118 take a float sample rate, the library takes the Rust crate's u32 (mod.rs:46).  Rounding a
    fractional or sub-1 Hz rate would silently change the phasor: the adapters refuse it before any GPU work."""
    from caf_cookoff_b200 import siblings
    x = np.ones(8, dtype=np.complex128)
    for bad in (47999.5, 0.25, 0.0, -48000.0, 2.0 ** 32, float("nan")):
        with pytest.raises(ValueError):
            siblings.PythonSibling.apply_fdoa(x, 1.0, bad)
        with pytest.raises(ValueError):
            siblings.PythonSibling.amb_surf(x, x, [0.0], bad)
    assert siblings._whole_sample_rate(48000.0) == 48000 and siblings._whole_sample_rate(1) == 1
