"""Development: phase stamps of the cluster-fused long-row kernel (needs the -DCAF_TRACE build).
   CAF_B200_SO=scripts/micro/libcaf_b200_trace.so python scripts/trace_cluster.py"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caf_cookoff_b200 import _lib, surface_arrays, default_handle, generate as G
needle, hay = G.as_inputs(G.pair(0, seed=0, chirp_length=int(os.environ.get("L", "32768"))))
freqs = np.linspace(-100, 100, 512, endpoint=False)
h = default_handle(); lib = _lib.load()
lib.caf_b200_debug_trace(h.raw, None, 0)
for _ in range(2): surface_arrays(needle, hay, freqs, 48000, want_surface=True)
ncta = 148
raw = np.zeros(ncta * 16 * 8 * 32, dtype=np.int64)
assert lib.caf_b200_debug_trace(h.raw, raw.ctypes.data_as(C.c_void_p), ncta) == 0
nslot = 148 * 2 * 8 * 16
buf = raw[:nslot].reshape(148, 2, 8, 16)
names = ["S compute", "wait a", "spread stores", "barrier b", "core", "(arrive c)+twiddle+ptab", "wait c", "gather stores", "barrier d", "gather dft+swap+emit", "argmax"]
print("median cycles per phase over CTAs, rows 1..6 of each cluster: group 0 | group 1")
for s_ in range(11):
    row = []
    for r in (0, 1):
        a_, b_ = buf[:, r, 1:7, s_], buf[:, r, 1:7, s_ + 1]
        ok = (a_ > 0) & (b_ > 0)
        row.append(float(np.median((b_ - a_)[ok])) if ok.any() else float("nan"))
    print(f"{s_:2d} {names[s_]:28s} {row[0]:8.0f} {row[1]:8.0f}")
a_, b_ = buf[:, 0, 1:6, 0], buf[:, 0, 2:7, 0]
ok = (a_ > 0) & (b_ > 0)
print("row period median", float(np.median((b_ - a_)[ok])), "active CTAs", int((buf[:, 0, 0, 0] > 0).sum()))
