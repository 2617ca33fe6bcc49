"""Sample-file I/O of the reference (caf_rust/src/utils.rs) and its doppler-grid helpers.

    read_file_c64      utils.rs:10-35   packed little-endian f32 I/Q  ->  complex128
    write_file_binary  utils.rs:39-63   complex128 -> raw little-endian f64 pairs (numpy complex128)
    gen_float_shifts   tests/test.rs:335-352   integer milli-Hz half-open range / 1e3
    bench_shifts       benches/caf_bench.rs:30-35, src/main.rs:19-22   400 rows, -100.0 .. 99.5 Hz
"""
from __future__ import annotations

import numpy as np


def read_file_c64(filename: str) -> np.ndarray:
    """Reads packed 32-bit floats, returns complex128 (Complex64 in Rust's naming, utils.rs:8-9).

    Like the reference, a trailing partial sample is an error (the Rust slice index panics)."""
    raw = np.fromfile(filename, dtype=np.uint8)
    if raw.size % 8:
        raise ValueError(f"{filename}: size {raw.size} is not a whole number of complex64 samples")
    return raw.view("<f4").astype(np.float64).view(np.complex128)


def write_file_binary(samples, filename: str) -> None:
    """BinaryIO::write_file_binary for Vec<Complex64>: numpy complex128 compatible."""
    np.ascontiguousarray(samples, dtype="<c16").tofile(filename)


def gen_float_shifts(start: float, end: float, step: float) -> np.ndarray:
    s, e, st = int(start * 1000.0), int(end * 1000.0), int(step * 1000.0)   # `as i32` / `as usize` truncate
    return np.array([m / 1e3 for m in range(s, e, st)], dtype=np.float64)


def bench_shifts() -> np.ndarray:
    return np.array([m / 1e3 for m in range(-100000, 100000, 500)], dtype=np.float64)
