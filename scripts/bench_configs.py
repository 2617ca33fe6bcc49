#!/usr/bin/env python
"""BASELINE configs 3, 4 and 5 (the non-headline shapes), one process per GPU.

    python scripts/bench_configs.py --config 3 [--rows 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_configs.py --config 3

  config 3: 4096 doppler x 65536 delay fp64 surface + peak, doppler ROWS sharded across ranks (strong scaling);
            the only collective is the packed-peak all_gather (32 B per rank) of caf_cookoff_b200.dist.exchange_peak.
  config 4: 4096 independent pairs of 400 x 8192, PAIRS sharded, peaks only.
  config 5: 16384 doppler x 2^20 delay, peak only (surface never materialised), rows sharded + packed-peak all_gather.
Prints one JSON line on rank 0.  Device-resident inputs, CUDA events on the launching stream, max over ranks.
"""
import argparse, json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5])
    ap.add_argument("--rows", type=int, default=0, help="doppler rows (configs 3/5) or pairs (config 4); 0 = the BASELINE size")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--exchange", choices=["abi", "torch"], default="abi",
                    help="peak exchange of configs 3/5: the library's own NCCL communicator (caf_b200_peak_allgather_dev) "
                         "or torch.distributed.all_gather_into_tensor")
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from caf_cookoff_b200 import Handle, _lib, generate as G, dist as cdist, bench_shifts, read_file_c64
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    h = Handle(local, stream=stream.cuda_stream); lib = _lib.load()
    FS = 48000
    comm = None
    if args.exchange == "abi" and args.config in (3, 5):
        import tempfile
        id_path = os.path.join(tempfile.gettempdir(), "caf_nccl_id_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x")))
        comm = cdist.Comm(h, world, rank, id_path)        # per-run nonce inside the file: a stale id is never picked up
    if args.config in (3, 5):
        L = 32768 if args.config == 3 else 1 << 19
        D = args.rows or (4096 if args.config == 3 else 16384)
        needle, hay = G.as_inputs(G.pair(0, seed=0, chirp_length=L))
        freqs = np.linspace(-100.0, 100.0, D, endpoint=False)
        lo, hi = cdist.shard_bounds(D, world, rank)
        want_surface = args.config == 3
        nd = torch.from_numpy(needle).to(dev); hd = torch.from_numpy(hay).to(dev)
        fd = torch.from_numpy(freqs[lo:hi].copy()).to(dev)
        d_loc = hi - lo
        surf = torch.empty((d_loc, 2 * L), dtype=torch.float64, device=dev) if want_surface else None
        rv = torch.empty(d_loc, dtype=torch.float64, device=dev); ri = torch.empty(d_loc, dtype=torch.int64, device=dev)
        pk = torch.zeros(4, dtype=torch.int64, device=dev)
        result = {}

        def step():
            rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), d_loc, FS,
                                            surf.data_ptr() if want_surface else None, rv.data_ptr(), ri.data_ptr(), pk.data_ptr())
            assert rc == 0, lib.caf_b200_last_error()
            if comm is not None:
                # pack on the device, ncclAllGather of 32 B per rank on the handle's stream, resolve: all inside the library
                result["peak"] = comm.peak_allgather_dev(pk.data_ptr(), lo)
                return
            w = pk.clone()                                   # [value bits, freq bits, doppler_idx, delay_idx]
            # pack: caf_b200_peak_pack layout = [value bits, global row, delay, freq bits]
            words = torch.stack([w[0], torch.where(w[2] == -1, w[2], w[2] + lo), w[3], w[1]])
            if world > 1:
                out = torch.empty(4 * world, dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(out, words)      # the one collective on this path: 32 bytes per rank
            else:
                out = words
            result["words"] = out
        cells = D * 2 * L
        what = f"cfg{args.config}: {D} doppler x {2*L} delay fp64, rows sharded x{world}, " + ("surface + peak" if want_surface else "peak only")
    else:
        P = args.rows or 4096
        data = os.path.join(ROOT, "tests", "golden", "data")
        names = sorted(os.listdir(data))
        ns = np.stack([read_file_c64(os.path.join(data, f"chirp_{i}_raw.c64")) for i in range(10)])
        hs = np.stack([read_file_c64(os.path.join(data, [n for n in names if n.startswith(f"chirp_{i}_T")][0]))[:4096] for i in range(10)])
        lo, hi = cdist.shard_bounds(P, world, rank)
        idx = np.arange(lo, hi) % 10
        L, D = 4096, 400
        freqs = bench_shifts()
        nd = torch.from_numpy(ns[idx]).to(dev); hd = torch.from_numpy(hs[idx]).to(dev); fd = torch.from_numpy(freqs).to(dev)
        p_loc = hi - lo
        pk = torch.zeros((p_loc, 4), dtype=torch.int64, device=dev)
        result = {}

        def step():
            rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), p_loc, L, fd.data_ptr(), D, FS,
                                            None, None, None, pk.data_ptr())
            assert rc == 0, lib.caf_b200_last_error()
            result["words"] = pk
        cells = P * D * 2 * L
        what = f"cfg4: {P} independent pairs of 400 x 8192 fp64, pairs sharded x{world}, peaks only"

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    line = {"workload": what, "n_gpus": world, "ms_per_step": ms, "cells_per_s": cells / (ms * 1e-3), "steps": args.steps}
    if args.config in (3, 5):
        from caf_cookoff_b200 import api
        if comm is not None:
            g = result["peak"]
            line["exchange"] = "caf_b200_peak_allgather_dev (library-owned NCCL communicator)"
        else:
            w = result["words"].cpu().numpy().view(np.uint64).reshape(-1, 4)
            g = api.peak_resolve(w)
            line["exchange"] = "torch.distributed all_gather_into_tensor"
        n_ = 2 * L
        line.update({"peak": {"freq_hz": g.freq_hz, "delay_idx": int(g.delay_idx), "doppler_idx": int(g.doppler_idx), "value": g.value},
                     "algorithmic_tflops": D * (10.0 * n_ * np.log2(n_) + 15.0 * n_) / (ms * 1e-3) / 1e12})
    else:
        w = result["words"].cpu().numpy()
        line["first_peaks"] = [[float(w[i].view(np.float64)[1]), int(w[i].view(np.uint64)[3])] for i in range(min(3, len(w)))]
        line["algorithmic_tflops"] = P * D * (10.0 * 8192 * 13 + 15.0 * 8192) / (ms * 1e-3) / 1e12
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
