"""Development A/B: time the surface kernel of one or more builds of the library.
   python scripts/quick_bench.py path/to/libA.so path/to/libB.so ...   (each build runs in its own process)
Prints per build: us per 400 x 8192 fp64 surface (L2 flushed, CUDA events, median of 200) and us per row in steady
state (148 pairs x 400 rows, peaks only), plus the peak of the chirp_0 pair as a sanity check."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def child():
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    sys.argv = ["bench.py"]
    import bench
    from caf_cookoff_b200 import Handle, _lib, bench_shifts
    f32 = bool(os.environ.get("QB_F32"))
    cdt = np.complex64 if f32 else np.complex128
    trdt = torch.float32 if f32 else torch.float64
    sfx = "f32" if f32 else "f64"
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
    needle, hay = bench.load_pair(0); freqs = bench_shifts(); D = freqs.size; L = 4096; N = 8192
    nd = torch.from_numpy(needle.astype(cdt)).to(dev); hd = torch.from_numpy(hay.astype(cdt)).to(dev)
    fd = torch.from_numpy(freqs).to(dev)
    surf = torch.empty((D, N), dtype=trdt, device=dev); rv = torch.empty(D, dtype=trdt, device=dev)
    ri = torch.empty(D, dtype=torch.int64, device=dev); pk = torch.zeros(4, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = getattr(lib, f"caf_b200_batch_{sfx}_dev")
    def step(P, n_, h_, s_, rv_, ri_, pk_):
        rc = fn(h.raw, n_.data_ptr(), h_.data_ptr(), P, L, fd.data_ptr(), D, 48000, s_, rv_.data_ptr(), ri_.data_ptr(), pk_.data_ptr())
        assert rc == 0, lib.caf_b200_last_error().decode()
    ts = []
    for i in range(220):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); step(1, nd, hd, surf.data_ptr(), rv, ri, pk); e1.record(stream)
        torch.cuda.synchronize()
        if i >= 20: ts.append(e0.elapsed_time(e1) * 1e3)
    p = pk.cpu().numpy(); peak = (float(p.view(np.float64)[1]), int(p.view(np.uint64)[3]))
    P = 148
    nb = nd.repeat(P, 1).contiguous(); hb = hd.repeat(P, 1).contiguous()
    rvb = torch.empty(P * D, dtype=trdt, device=dev); rib = torch.empty(P * D, dtype=torch.int64, device=dev)
    pkb = torch.zeros(4 * P, dtype=torch.int64, device=dev)
    tb = []
    for i in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); step(P, nb, hb, 0, rvb, rib, pkb); e1.record(stream)
        torch.cuda.synchronize()
        if i >= 1: tb.append(e0.elapsed_time(e1) * 1e3)
    pb = pkb.cpu().numpy().reshape(P, 4); okb = all(float(q.view(np.float64)[1]) == peak[0] and int(q.view(np.uint64)[3]) == peak[1] for q in pb)
    cyc_row = np.median(tb) / D * 1.965e3
    print(f"{os.environ.get('CAF_B200_SO', 'default'):50s} surface mean {np.mean(ts):6.2f} us (median {np.median(ts):6.2f}, min {np.min(ts):6.2f}; events tick at ~1 us)  steady {np.median(tb)/D:6.3f} us/row = {cyc_row:6.0f} cyc  peak {peak} batch_ok {okb}", flush=True)

if __name__ == "__main__":
    if os.environ.get("QB_CHILD"):
        child()
    else:
        for so in sys.argv[1:] or [""]:
            env = dict(os.environ, QB_CHILD="1")
            if so: env["CAF_B200_SO"] = os.path.abspath(so)
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
