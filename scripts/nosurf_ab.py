import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.argv=["bench.py"]
import bench
from caf_cookoff_b200 import Handle, _lib, bench_shifts
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
needle, hay = bench.load_pair(0); freqs = bench_shifts(); D = freqs.size; L=4096; N=8192
nd = torch.from_numpy(needle).to(dev); hd = torch.from_numpy(hay).to(dev); fd = torch.from_numpy(freqs).to(dev)
surf = torch.empty((D, N), dtype=torch.float64, device=dev); rv = torch.empty(D, dtype=torch.float64, device=dev)
ri = torch.empty(D, dtype=torch.int64, device=dev); pk = torch.zeros(4, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, sp in (("with surface", surf.data_ptr()), ("peak only", 0), ("with surface", surf.data_ptr()), ("peak only", 0)):
    ts=[]
    for i in range(120):
        flush.zero_(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(stream); rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), D, 48000, sp, rv.data_ptr(), ri.data_ptr(), pk.data_ptr()); e1.record(stream)
        torch.cuda.synchronize(); assert rc == 0
        if i >= 20: ts.append(e0.elapsed_time(e1)*1e3)
    print(name, "mean %.2f us median %.2f" % (np.mean(ts), np.median(ts)))
