"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_check.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caf_cookoff_b200 import (CafB200, Xcor, batch_arrays, generate as G, read_file_c64, surface_arrays, surface_layout, api)
D = os.path.join(ROOT, "tests/golden/data/")
needle = read_file_c64(D + "chirp_0_raw.c64"); hay = read_file_c64(D + "chirp_0_T+202samp_F+69.25Hz.c64")[:4096]
f = np.array([68.5, 69.0, 69.25, 70.0, -3.0])
s, pi, pv, pk = surface_arrays(needle, hay, f, 48000); print("rows", pk.freq_hz, pk.delay_idx)
surface_arrays(needle, hay, np.linspace(-100, 100, 300, endpoint=False), 48000)           # pipelined host path, 2-3 rows per CTA
surface_arrays(needle[:1000], hay[:1000], f, 48000)
surface_arrays(needle, hay, f, 48000, variant=api._Variant32)
batch_arrays(np.stack([needle, needle]), np.stack([hay, hay]), f, 48000)
CafB200.apply_freq_shift(needle[:777], 12.5, 48000)
Xcor.new(8192).run(np.concatenate([needle, needle]), np.concatenate([hay, hay])); Xcor.new(1000).run(needle[:1000], hay[:1000])
surface_layout(needle, hay, f, 48000, 1); surface_layout(needle, hay, f, 48000, 2)
for l in (5000, 20001):                                                                   # one level, R = 2 and 8
    n2, h2 = G.as_inputs(G.pair(0, seed=0, chirp_length=l))
    surface_arrays(n2, h2, f[:3], 48000)
n3, h3 = G.as_inputs(G.pair(0, seed=0, chirp_length=100000))                              # two levels, fused kernels
_, pi3, _, pk3 = surface_arrays(n3, h3, f[:2], 48000, want_surface=False); print("two-level", pk3.freq_hz, pk3.delay_idx)
# round 2 paths: device-resident surface object with lazy rows, peak-only call (pinned completion word, spin wait),
# inputs in one block / scattered, the sharded entry over a world-1 communicator (pack in the kernel, device resolve)
import ctypes as C, tempfile
from caf_cookoff_b200 import _lib, default_handle
from caf_cookoff_b200.dist import Comm
lib = _lib.load(); h = default_handle()
so = C.c_void_p(); pk = _lib.Peak()
assert lib.caf_b200_surface_create_f64(h.raw, needle.ctypes.data, hay.ctypes.data, 4096, f.ctypes.data, f.size, 48000, C.byref(so)) == 0
row = np.empty(8192); assert lib.caf_b200_surface_fetch_rows(so, 1, 1, row.ctypes.data) == 0
pv = np.empty(f.size); pi_ = np.empty(f.size, dtype=np.uint64); fz = np.empty(f.size)
assert lib.caf_b200_surface_row_peaks(so, fz.ctypes.data, pv.ctypes.data, pi_.ctypes.data) == 0
assert lib.caf_b200_surface_find_peak(so, C.byref(pk)) == 0 and row[int(pi_[1])] == pv[1]
lib.caf_b200_surface_destroy(so)
assert lib.caf_b200_peak_f64(h.raw, needle.ctypes.data, hay.ctypes.data, 4096, f.ctypes.data, f.size, 48000, C.byref(pk)) == 0
print("surface object / peak-only", pk.freq_hz, pk.delay_idx)
import torch
comm = Comm(h, 1, 0, os.path.join(tempfile.gettempdir(), "caf_sanitize_id_%d" % os.getpid()))
nd = torch.from_numpy(needle).cuda(); hd = torch.from_numpy(hay).cuda(); fd = torch.from_numpy(f).cuda(); outp = torch.zeros(4, dtype=torch.int64, device="cuda")
comm.sharded_dev(nd.data_ptr(), hd.data_ptr(), 4096, fd.data_ptr(), f.size, 0, 48000, outp.data_ptr()); h.sync()
n3d = torch.from_numpy(n3).cuda(); h3d = torch.from_numpy(h3).cuda()
comm.sharded_dev(n3d.data_ptr(), h3d.data_ptr(), n3.size, fd.data_ptr(), 2, 0, 48000, outp.data_ptr()); h.sync()
comm.close()
print("sanitize pass done")
