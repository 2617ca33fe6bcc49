"""Development: where the ~70 us of the peak-only host call go (wall clock per call, no flush, 2000 calls each)."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.argv = ["bench.py"]
import bench
from caf_cookoff_b200 import Handle, _lib, bench_shifts
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
needle, hay = bench.load_pair(0); freqs = bench_shifts(); D = freqs.size; L = 4096
nh = torch.from_numpy(needle).pin_memory(); hh = torch.from_numpy(hay).pin_memory(); fh = torch.from_numpy(freqs).pin_memory()
nd = nh.to(dev); hd = hh.to(dev); fd = fh.to(dev)
pk = _lib.Peak(); pkd = torch.zeros(4, dtype=torch.int64, device=dev)
def t(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def host_peak():
    rc = lib.caf_b200_peak_f64(h.raw, nh.data_ptr(), hh.data_ptr(), L, fh.data_ptr(), D, 48000, C.cast(C.byref(pk), C.c_void_p)); assert rc == 0
def dev_sync():
    rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), D, 48000, None, None, None, pkd.data_ptr()); assert rc == 0
    lib.caf_b200_sync(h.raw)
def dev_nosync():
    rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), D, 48000, None, None, None, pkd.data_ptr()); assert rc == 0
def copies_sync():
    nd.copy_(nh, non_blocking=True); hd.copy_(hh, non_blocking=True); fd.copy_(fh, non_blocking=True); torch.cuda.synchronize()
def sync_only():
    lib.caf_b200_sync(h.raw)
print(f"host peak call (3 H2D + kernel + zero-copy peak + sync): {t(host_peak):7.2f} us")
print(f"device call + sync (inputs resident):                    {t(dev_sync):7.2f} us")
print(f"device call, no sync (back to back, throughput):         {t(dev_nosync):7.2f} us")
print(f"three H2D copies (torch) + sync:                         {t(copies_sync):7.2f} us")
print(f"sync of an idle stream:                                  {t(sync_only):7.2f} us")
