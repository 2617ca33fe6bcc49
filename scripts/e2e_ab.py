import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.argv=["bench.py"]
import bench
from caf_cookoff_b200 import Handle, _lib, bench_shifts
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
needle, hay = bench.load_pair(0); freqs = bench_shifts(); D = freqs.size; L=4096; N=8192
nh = torch.from_numpy(needle).pin_memory(); hh = torch.from_numpy(hay).pin_memory(); fh = torch.from_numpy(freqs).pin_memory()
sh = torch.empty((D, N), dtype=torch.float64).pin_memory(); rv = torch.empty(D, dtype=torch.float64).pin_memory(); ri = torch.empty(D, dtype=torch.int64).pin_memory()
pk = _lib.Peak(); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    rc = lib.caf_b200_surface_f64(h.raw, nh.data_ptr(), hh.data_ptr(), L, fh.data_ptr(), D, 48000, sh.data_ptr(), rv.data_ptr(), ri.data_ptr(), C.cast(C.byref(pk), C.c_void_p)); assert rc == 0
for _ in range(5): step()
ts=[]
for i in range(100):
    flush.zero_(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(stream); step(); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
print("zerocopy env", os.environ.get("CAF_B200_ZEROCOPY"), "e2e us median %.1f min %.1f" % (np.median(ts), np.min(ts)), "peak", pk.freq_hz, pk.delay_idx, "sum", float(sh.sum()))
