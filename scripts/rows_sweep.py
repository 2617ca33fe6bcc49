"""Development: surface-kernel time against the number of doppler rows (1, 2, 3 rows per CTA on 148 SMs), L2 flushed.
   python scripts/rows_sweep.py          -> fixed cost per launch = T(148 rows) - one steady-state row"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
sys.argv = ["bench.py"]
import bench
from caf_cookoff_b200 import Handle, _lib

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
needle, hay = bench.load_pair(0)
L, N = 4096, 8192
nd = torch.from_numpy(needle).to(dev); hd = torch.from_numpy(hay).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for D in (2, 148, 296, 400, 444):
    freqs = np.linspace(-100.0, 100.0, D, endpoint=False)
    fd = torch.from_numpy(freqs).to(dev)
    surf = torch.empty((D, N), dtype=torch.float64, device=dev); rv = torch.empty(D, dtype=torch.float64, device=dev)
    ri = torch.empty(D, dtype=torch.int64, device=dev); pk = torch.zeros(4, dtype=torch.int64, device=dev)
    for want_surface, want_peak, do_flush in ((True, True, True), (False, True, True), (True, False, True), (True, True, False)):
        ts = []
        for i in range(120):
            if do_flush: flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), D, 48000,
                                            surf.data_ptr() if want_surface else 0, rv.data_ptr(), ri.data_ptr(), pk.data_ptr() if want_peak else 0)
            e1.record(stream)
            assert rc == 0, lib.caf_b200_last_error().decode()
            torch.cuda.synchronize()
            if i >= 20: ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"D = {D:5d} ({D / 148:5.2f} rows per SM) surface={int(want_surface)} fused_peak={int(want_peak)} l2_flush={int(do_flush)}: median {np.median(ts):7.2f} us  min {np.min(ts):7.2f} us", flush=True)
# an empty launch between the same events: the floor of the measurement itself
ts = []
x = torch.zeros(1, device=dev)
for i in range(120):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream); x.add_(1.0); e1.record(stream); torch.cuda.synchronize()
    if i >= 20: ts.append(e0.elapsed_time(e1) * 1e3)
print(f"one-element torch kernel between the same events: median {np.median(ts):.2f} us", flush=True)

# with the -DCAF_TRACE build (CAF_B200_SO=scripts/micro/libcaf_b200_trace.so): where the launch's time sits relative
# to the CTAs' own lifetimes (global timer): entry of the first CTA ... end of the row work ... end of the find_peak tail
import ctypes as C
if lib.caf_b200_debug_trace(h.raw, None, 0) == 0:
    D = 400
    freqs = np.linspace(-100.0, 100.0, D, endpoint=False); fd = torch.from_numpy(freqs).to(dev)
    surf = torch.empty((D, N), dtype=torch.float64, device=dev); rv = torch.empty(D, dtype=torch.float64, device=dev)
    ri = torch.empty(D, dtype=torch.int64, device=dev); pk = torch.zeros(4, dtype=torch.int64, device=dev)
    buf = np.zeros((148, 16, 8, 32), dtype=np.int64)
    for i in range(8):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), D, 48000,
                                        surf.data_ptr(), rv.data_ptr(), ri.data_ptr(), pk.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        ev_us = e0.elapsed_time(e1) * 1e3
        if lib.caf_b200_debug_trace(h.raw, buf.ctypes.data_as(C.c_void_p), 148) != 0: break
        g0 = buf[:, 0, 0, 20]; g1 = buf[:, 0, 0, 21]; g2 = buf[:, 0, 0, 29]; c0 = buf[:, 0, 0, 22]; c1 = buf[:, 0, 0, 23]
        ok = g0 > 0
        if not ok.any(): print("trace build not loaded (no stamps)"); break
        base = g0[ok].min()
        print(f"trace D=400: events {ev_us:6.2f} us | first CTA entry 0, last CTA entry {(g0[ok].max() - base) / 1e3:5.2f}, "
              f"last row work done {(g1[ok].max() - base) / 1e3:6.2f}, last CTA exit (after find_peak tail) {(g2[ok].max() - base) / 1e3:6.2f} us"
              f" | SM clock over CTA lifetimes {np.median((c1 - c0)[ok] / np.maximum(g1 - g0, 1)[ok]):.3f} GHz", flush=True)
