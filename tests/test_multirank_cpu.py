"""World-size-2 gloo test of the N > 1 path on CPU: row sharding + packed-peak exchange == unsharded find_peak.

Each rank gets its shard's peak from the CPU oracle (there is no GPU here, and the product path has no CPU
fallback); what is under test is the host logic around it: shard bounds, caf_b200_peak_pack, the all_gather,
caf_b200_peak_resolve with the reference's tie-break."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from caf_cookoff_b200 import _lib, dist as cdist
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = os.path.join(ROOT, "tests", "golden", "data")
        needle = O.read_file_c64(os.path.join(data, "chirp_9_raw.c64"))
        hay = O.read_file_c64(os.path.join(data, "chirp_9_T+176samp_F+61.49Hz.c64"))[:4096]
        shifts = O.gen_float_shifts(40.0, 80.0, 0.5)                  # 80 rows, answer 61.5 Hz is in rank 1's shard
        lo, hi = cdist.shard_bounds(len(shifts), world, rank)
        _, pidx, pval = O.caf_surface(needle, hay, shifts[lo:hi], 48000, want_surface=False)
        best = int(np.argmax(pval))
        local = _lib.Peak(float(pval[best]), float(shifts[lo + best]), best, int(pidx[best]))
        out = cdist.exchange_peak(local, lo)
        q.put((rank, out.value, out.freq_hz, int(out.doppler_idx), int(out.delay_idx), lo, hi))
    finally:
        dist.destroy_process_group()


def test_row_sharded_peak_exchange_world2():
    import torch.multiprocessing as mp
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = os.path.join(ROOT, "tests", "golden", "data")
    needle = O.read_file_c64(os.path.join(data, "chirp_9_raw.c64"))
    hay = O.read_file_c64(os.path.join(data, "chirp_9_T+176samp_F+61.49Hz.c64"))[:4096]
    shifts = O.gen_float_shifts(40.0, 80.0, 0.5)
    _, pidx, pval = O.caf_surface(needle, hay, shifts, 48000, want_surface=False)
    want_f, want_idx = O.find_peak(shifts, pidx, pval)
    assert (want_f, want_idx) == (61.5, 176)
    assert res[0][5:] == (0, 40) and res[1][5:] == (40, 80)
    for r in res:                                   # every rank resolves the same global winner
        assert r[2] == want_f and r[4] == want_idx
        assert r[3] == int(np.argmax(pval)) and r[1] == float(pval.max())


def test_shard_bounds_cover_everything():
    from caf_cookoff_b200.dist import shard_bounds
    for n in (0, 1, 7, 400, 4096):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
