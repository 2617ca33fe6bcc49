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
from .api import peak_pack, peak_resolve


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
