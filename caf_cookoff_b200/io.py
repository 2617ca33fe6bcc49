"""Sample-file I/O of the reference (caf_rust/src/utils.rs) and its doppler-grid helpers.

    read_file_c64      utils.rs:10-35   packed little-endian f32 I/Q  ->  complex128
    read_file_c64_dev  the same samples, loaded through pinned memory straight onto the GPU (caf_b200_load_c64_dev_*)
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


class DeviceSamples:
    """Samples resident on the GPU (what `read_file_c64_dev` returns): `.ptr` goes to the *_dev entry points."""

    def __init__(self, handle, ptr: int, n: int, f32: bool):
        self.handle, self.ptr, self.size, self.f32 = handle, ptr, n, f32

    @property
    def dtype(self):
        return np.complex64 if self.f32 else np.complex128

    def to_host(self) -> np.ndarray:
        from . import _lib
        from .api import _check
        out = np.empty(self.size, dtype=self.dtype)
        if self.size:
            _check(_lib.load().caf_b200_dev_download(self.handle.raw, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            from . import _lib
            _lib.load().caf_b200_dev_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def read_file_c64_dev(filename: str, first_sample: int = 0, max_samples: int = 0, *, f32: bool = False, handle=None) -> DeviceSamples:
    """`read_file_c64` (utils.rs:10-35) straight onto the GPU: the file's packed f32 pairs are read into pinned memory,
    cross PCIe as they are (8 bytes per sample) and are widened on the device -- bit-identical to the host loader's
    samples.  `first_sample` / `max_samples` select a window (main.rs:15 truncates the haystack to the needle's length)."""
    import ctypes as C
    from . import _lib
    from .api import _check, default_handle
    h = handle or default_handle()
    ptr, n = C.c_void_p(), C.c_size_t()
    fn = getattr(_lib.load(), "caf_b200_load_c64_dev_" + ("f32" if f32 else "f64"))
    _check(fn(h.raw, str(filename).encode(), int(first_sample), int(max_samples), C.byref(ptr), C.byref(n)))
    return DeviceSamples(h, ptr.value or 0, int(n.value), f32)


def write_file_binary(samples, filename: str) -> None:
    """BinaryIO::write_file_binary for Vec<Complex64>: numpy complex128 compatible."""
    np.ascontiguousarray(samples, dtype="<c16").tofile(filename)


def gen_float_shifts(start: float, end: float, step: float) -> np.ndarray:
    s, e, st = int(start * 1000.0), int(end * 1000.0), int(step * 1000.0)   # `as i32` / `as usize` truncate
    return np.array([m / 1e3 for m in range(s, e, st)], dtype=np.float64)


def bench_shifts() -> np.ndarray:
    return np.array([m / 1e3 for m in range(-100000, 100000, 500)], dtype=np.float64)
