"""Quick on-GPU sanity pass used during development (not a test): parity of every entry point."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caf_cookoff_b200 import (CafB200, CafB200F32, Xcor, XcorF32, read_file_c64, bench_shifts, surface_arrays,
                              batch_arrays, api)
from oracle import oracle as O

D = os.path.join(ROOT, "tests/golden/data/")
needle = read_file_c64(D + "chirp_0_raw.c64")
hay = read_file_c64(D + "chirp_0_T+202samp_F+69.25Hz.c64")[:4096]
sh = bench_shifts()

def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())

t = time.time(); surf, pidx, pval, pk = surface_arrays(needle, hay, sh, 48000); t1 = time.time() - t
osurf, opidx, opval = O.caf_surface(needle, hay, sh, 48000)
print("f64 surface rel", rel(surf, osurf), "pidx eq", (pidx == opidx).all(), "pval rel", rel(pval, opval),
      "peak", pk.freq_hz, pk.delay_idx, pk.doppler_idx, pk.value, "oracle", O.find_peak(sh, opidx, opval), f"{t1*1e3:.1f} ms first call")
t = time.time(); surface_arrays(needle, hay, sh, 48000); print("second call ms", (time.time() - t) * 1e3)
t = time.time(); surface_arrays(needle, hay, sh, 48000, want_surface=False); print("peak-only call ms", (time.time() - t) * 1e3)

s32, pidx32, pval32, pk32 = surface_arrays(needle, hay, sh, 48000, variant=api._Variant32)
print("f32 surface rel", rel(s32.astype(np.float64), osurf), "pidx eq", (pidx32 == opidx).mean(), "peak", pk32.freq_hz, pk32.delay_idx)

x = needle[:1000]
print("shift rel", rel(CafB200.apply_freq_shift(x, 77.77, 48000), O.apply_freq_shift(x, 77.77, 48000)))
for n in (8192, 4096, 1000, 37, 1):
    rng = np.random.default_rng(n)
    a = rng.normal(size=n) + 1j * rng.normal(size=n); b = rng.normal(size=n) + 1j * rng.normal(size=n)
    print("xcor n", n, rel(Xcor.new(n).run(a, b), O.xcor(a, b)))
for L in (4095, 2048, 1000, 17, 1):
    f = sh[::50]
    s, pi, pv, p = surface_arrays(needle[:L], hay[:L], f, 48000)
    os_, opi, opv = O.caf_surface(needle[:L], hay[:L], f, 48000)
    print("L", L, "rel", rel(s, os_), "pidx eq", (pi == opi).all(), "peak", (p.freq_hz, p.delay_idx), O.find_peak(f, opi, opv))
ns = np.stack([read_file_c64(D + f"chirp_{i}_raw.c64") for i in range(3)])
hs = np.stack([read_file_c64(D + n)[:4096] for n in ("chirp_0_T+202samp_F+69.25Hz.c64", "chirp_1_T+78samp_F+35.99Hz.c64", "chirp_2_T+169samp_F+32.16Hz.c64")])
s, pi, pv, pks = batch_arrays(ns, hs, sh, 48000, want_surface=True)
for i in range(3):
    os_, opi, opv = O.caf_surface(ns[i], hs[i], sh, 48000)
    print("batch", i, rel(s[i], os_), (pi[i] == opi).all(), (pks[i].freq_hz, pks[i].delay_idx), O.find_peak(sh, opi, opv))
