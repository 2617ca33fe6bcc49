"""Multi-GPU plumbing for the CAF path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards without any data-path collective: doppler rows are independent
(mod.rs:185,283 — the reference's own par_iter) and so are signal pairs.  The only cross-rank step is find_peak's
maximum (mod.rs:36-40): each rank packs its local peak into 4 uint64 words (caf_b200_peak_pack), ONE all_gather of
32 bytes per rank moves them, and every rank resolves the same winner with the reference's tie-break
(caf_b200_peak_resolve: larger value, then lower global doppler row).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _lib
from .api import _check, peak_pack, peak_resolve


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of the n rows (or pairs) owned by `rank`: [lo, hi)."""
    return n * rank // world, n * (rank + 1) // world


def exchange_peak(local: _lib.Peak, global_row_offset: int, device=None, group=None) -> _lib.Peak:
    """All ranks call this with their shard's peak; all ranks get the global find_peak result."""
    import torch
    import torch.distributed as dist
    words = peak_pack(local, global_row_offset)                       # uint64[4]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return peak_resolve(words.reshape(1, 4))
    world = dist.get_world_size(group)
    t = torch.from_numpy(words.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * 4, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return peak_resolve(out.cpu().numpy().view(np.uint64).reshape(world, 4))


def gather_pair_peaks(local_peaks, n_pairs: int, device=None, group=None) -> np.ndarray:
    """Pairs sharded across ranks (BASELINE config 4): rank r holds the peaks of pairs shard_bounds(n_pairs, world, r),
    as an array of caf_b200_peak records viewed as uint64 [p_local, 4] (value bits, freq bits, doppler_idx,
    delay_idx).  ONE all_gather of the padded blocks (32 bytes per pair) gives every rank all n_pairs records in pair
    order.  Pairs are independent, so unlike exchange_peak nothing is resolved: this is a pure gather."""
    import torch
    import torch.distributed as dist
    loc = np.ascontiguousarray(local_peaks).view(np.uint64).reshape(-1, 4)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if loc.shape[0] != n_pairs:
            raise ValueError(f"one rank must hold all {n_pairs} pairs, got {loc.shape[0]}")
        return loc.copy()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_pairs, world, rank)
    if loc.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns pairs [{lo}, {hi}) but passed {loc.shape[0]} peaks")
    width = max(b - a for a, b in (shard_bounds(n_pairs, world, r) for r in range(world)))
    block = np.zeros((width, 4), dtype=np.uint64)
    block[: hi - lo] = loc
    t = torch.from_numpy(block.view(np.int64))
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * width, 4), dtype=torch.int64, device=t.device)      # blocks concatenated along dim 0
    dist.all_gather_into_tensor(out, t, group=group)
    allw = out.cpu().numpy().view(np.uint64).reshape(world, width, 4)
    parts = []
    for r in range(world):
        a, b = shard_bounds(n_pairs, world, r)
        parts.append(allw[r, : b - a])
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 4), dtype=np.uint64)


def gather_surface(local_rows, n_rows: int, group=None):
    """The full surface, only when asked for (the path itself never needs it): rank r holds rows
    shard_bounds(n_rows, world, r) of the doppler grid as [hi - lo, 2L] (a numpy array, or a torch tensor — a CUDA
    tensor under NCCL, so the rows travel GPU to GPU over NVLink); ONE all_gather of the padded blocks gives every rank
    the [n_rows, 2L] surface in freqs_hz order, exactly the Vec<CafSurfaceRow> order of the ordered strategies
    (mod.rs:135-162).  Uneven shards are padded to the widest block and trimmed after the exchange."""
    import torch
    import torch.distributed as dist
    is_np = isinstance(local_rows, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local_rows)) if is_np else local_rows.contiguous()
    if t.dim() != 2:
        raise ValueError("local_rows must be [rows, 2L]")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if t.shape[0] != n_rows:
            raise ValueError(f"one rank must hold all {n_rows} rows, got {t.shape[0]}")
        return local_rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_rows, world, rank)
    if t.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns rows [{lo}, {hi}) but passed {t.shape[0]}")
    bounds = [shard_bounds(n_rows, world, r) for r in range(world)]
    width = max(b - a for a, b in bounds)
    cols = t.shape[1]
    if hi - lo == width:
        block = t
    else:
        block = torch.zeros((width, cols), dtype=t.dtype, device=t.device)
        block[: hi - lo] = t
    out = torch.empty((world * width, cols), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, block, group=group)
    if all(b - a == width for a, b in bounds):
        full = out
    else:
        full = torch.cat([out[r * width: r * width + (b - a)] for r, (a, b) in enumerate(bounds)], dim=0)
    return full.numpy() if is_np else full


def peaks_as_tuples(words: np.ndarray):
    """[(freq_hz, delay_idx)] per pair from gather_pair_peaks' uint64 records: what find_peak returns (mod.rs:31-42)."""
    w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, 4)
    return [(float(w[i, 1:2].view(np.float64)[0]), int(w[i, 3])) for i in range(w.shape[0])]


class Comm:
    """The library's own NCCL communicator (caf_b200_comm_*): what a caller without torch.distributed (the Rust
    shim) uses.  The 128-byte NCCL id travels out of band; here through a small file any rank can read."""

    def __init__(self, handle, world: int, rank: int, id_path: str, timeout_s: float = 120.0):
        import ctypes as C
        import os
        import time
        self._lib = _lib.load()
        self.world, self.rank, self.handle = world, rank, handle
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _check(self._lib.caf_b200_comm_unique_id(C.cast(buf, C.c_void_p)))
            tmp = id_path + ".tmp"
            with open(tmp, "wb") as f:
                f.write(bytes(buf))
            os.replace(tmp, id_path)                      # atomic publish
        else:
            t_end = time.time() + timeout_s
            while not os.path.exists(id_path):
                if time.time() > t_end:
                    raise TimeoutError(f"rank {rank}: no NCCL id at {id_path}")
                time.sleep(0.01)
            C.memmove(buf, open(id_path, "rb").read(128), 128)
        self._c = C.c_void_p()
        _check(self._lib.caf_b200_comm_create(handle.raw, world, rank, C.cast(buf, C.c_void_p), C.byref(self._c)))

    @property
    def raw(self):
        return self._c

    def shard(self, n: int) -> Tuple[int, int]:
        import ctypes as C
        lo, hi = C.c_size_t(), C.c_size_t()
        _check(self._lib.caf_b200_comm_shard(self._c, n, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def peak_allgather_dev(self, local_peak_dev_ptr: int, global_row_offset: int) -> _lib.Peak:
        import ctypes as C
        out = _lib.Peak()
        _check(self._lib.caf_b200_peak_allgather_dev(self.handle.raw, self._c, C.c_void_p(local_peak_dev_ptr),
                                                     int(global_row_offset), C.byref(out)))
        return out

    def surface_sharded(self, needle, haystack, freqs_hz, fs: int, want_surface: bool = True):
        """caf_b200_surface_sharded_f64: (local rows [hi-lo, 2L] or None, global Peak)."""
        import ctypes as C
        n_ = np.ascontiguousarray(needle, dtype=np.complex128).ravel()
        h_ = np.ascontiguousarray(haystack, dtype=np.complex128).ravel()
        f_ = np.ascontiguousarray(freqs_hz, dtype=np.float64).ravel()
        lo, hi = self.shard(f_.size)
        surf = np.empty((hi - lo, 2 * n_.size), dtype=np.float64) if want_surface else None
        pk = _lib.Peak()
        _check(self._lib.caf_b200_surface_sharded_f64(
            self.handle.raw, self._c, n_.ctypes.data_as(C.c_void_p), h_.ctypes.data_as(C.c_void_p), n_.size,
            f_.ctypes.data_as(C.c_void_p), f_.size, int(fs), None if surf is None else surf.ctypes.data_as(C.c_void_p),
            C.byref(pk)))
        return surf, pk

    def close(self):
        if self._c:
            self._lib.caf_b200_comm_destroy(self._c)
            import ctypes as C
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
