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


def _pair_worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from caf_cookoff_b200 import dist as cdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = cdist.shard_bounds(n_pairs, world, rank)
        local = _pair_peak_words(range(lo, hi))
        allw = cdist.gather_pair_peaks(local, n_pairs)
        q.put((rank, lo, hi, allw.tobytes()))
    finally:
        dist.destroy_process_group()


def _pair_peak_words(indices):
    """caf_b200_peak records (value, freq_hz, doppler_idx, delay_idx as 4 x 8 bytes) of the seed-0 chirp pairs,
    from the CPU oracle (the N > 1 host logic is what is under test; there is no GPU here)."""
    from oracle import oracle as O
    data = os.path.join(ROOT, "tests", "golden", "data")
    names = sorted(os.listdir(data))
    shifts = O.gen_float_shifts(-100.0, 100.0, 5.0)
    rows = []
    for i in indices:
        needle = O.read_file_c64(os.path.join(data, f"chirp_{i}_raw.c64"))
        hay = O.read_file_c64(os.path.join(data, [n for n in names if n.startswith(f"chirp_{i}_T")][0]))[:needle.size]
        _, pidx, pval = O.caf_surface(needle, hay, shifts, 48000, want_surface=False)
        best = int(np.argmax(pval))
        rec = np.zeros(4, dtype=np.uint64)
        rec[0:1] = np.array([pval[best]]).view(np.uint64)
        rec[1:2] = np.array([shifts[best]]).view(np.uint64)
        rec[2], rec[3] = best, int(pidx[best])
        rows.append(rec)
    return np.stack(rows) if rows else np.zeros((0, 4), dtype=np.uint64)


@pytest.mark.parametrize("world,n_pairs", [(2, 10), (3, 7)])
def test_pair_sharded_peaks_gather(world, n_pairs):
    """BASELINE config 4's N > 1 path: pairs sharded (evenly and unevenly), one all_gather of 32-byte records, every
    rank ends with all peaks in pair order — identical to the unsharded computation."""
    import torch.multiprocessing as mp
    from caf_cookoff_b200 import dist as cdist
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_pair_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _pair_peak_words(range(n_pairs))
    assert [r[1:3] for r in res] == [cdist.shard_bounds(n_pairs, world, r) for r in range(world)]
    for r in res:
        got = np.frombuffer(r[3], dtype=np.uint64).reshape(-1, 4)
        assert np.array_equal(got, want)
    tuples = cdist.peaks_as_tuples(want)
    assert len(tuples) == n_pairs and tuples[0][1] == 202        # chirp_0: delay 202 (test.rs)
    # world size 1 (no process group): the local block is the answer
    assert np.array_equal(cdist.gather_pair_peaks(want, n_pairs), want)
    with pytest.raises(ValueError):
        cdist.gather_pair_peaks(want[:-1], n_pairs)


def _surface_worker(rank, world, port, n_rows, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from caf_cookoff_b200 import dist as cdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        needle, hay, shifts = _small_case(n_rows)
        lo, hi = cdist.shard_bounds(n_rows, world, rank)
        from oracle import oracle as O
        local, _, _ = O.caf_surface(needle, hay, shifts[lo:hi], 48000)
        local = np.asarray(local).reshape(hi - lo, 2 * needle.size)
        full_np = cdist.gather_surface(local, n_rows)
        full_t = cdist.gather_surface(torch.from_numpy(local.copy()), n_rows)          # tensor in, tensor out
        assert isinstance(full_np, np.ndarray) and isinstance(full_t, torch.Tensor)
        assert np.array_equal(full_t.numpy(), full_np)
        q.put((rank, full_np.tobytes()))
    finally:
        dist.destroy_process_group()


def _small_case(n_rows):
    from oracle import oracle as O
    data = os.path.join(ROOT, "tests", "golden", "data")
    needle = O.read_file_c64(os.path.join(data, "chirp_0_raw.c64"))[:512]
    hay = O.read_file_c64(os.path.join(data, "chirp_0_T+202samp_F+69.25Hz.c64"))[:512]
    shifts = np.linspace(60.0, 75.0, n_rows)
    return needle, hay, shifts


@pytest.mark.parametrize("world,n_rows", [(2, 8), (3, 7)])
def test_row_sharded_surface_gather(world, n_rows):
    """'The full surface is gathered only when requested': rows sharded evenly and unevenly, one all_gather of the
    padded blocks, every rank ends with the unsharded surface bit for bit, in freqs_hz order."""
    import torch.multiprocessing as mp
    from caf_cookoff_b200 import dist as cdist
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30700 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_surface_worker, args=(r, world, port, n_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    needle, hay, shifts = _small_case(n_rows)
    want, _, _ = O.caf_surface(needle, hay, shifts, 48000)
    want = np.asarray(want).reshape(n_rows, 1024)
    for r in res:
        assert np.array_equal(np.frombuffer(r[1], dtype=np.float64).reshape(n_rows, 1024), want)
    # no process group: the local block is the surface
    assert cdist.gather_surface(want, n_rows) is want
    with pytest.raises(ValueError):
        cdist.gather_surface(want[:-1], n_rows)
