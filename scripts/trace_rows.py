"""Development: dump the per-warp phase timeline of the row kernel (needs the -DCAF_TRACE build).
   CAF_B200_SO=scripts/micro/libcaf_b200_trace.so python scripts/trace_rows.py"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caf_cookoff_b200 import _lib, read_file_c64, bench_shifts, surface_arrays, default_handle, api
VARIANT = api._Variant32 if os.environ.get("CAF_TRACE_F32") else api._Variant
D = os.path.join(ROOT, "tests/golden/data/")
needle = read_file_c64(D + "chirp_0_raw.c64"); hay = read_file_c64(D + "chirp_0_T+202samp_F+69.25Hz.c64")[:4096]
sh = bench_shifts()
h = default_handle(); lib = _lib.load()
lib.caf_b200_debug_trace(h.raw, None, 0)            # allocate
for _ in range(3): surface_arrays(needle, hay, sh, 48000, want_surface=True, variant=VARIANT)
ncta = 148
buf = np.zeros((ncta, 16, 8, 32), dtype=np.int64)
assert lib.caf_b200_debug_trace(h.raw, buf.ctypes.data_as(C.c_void_p), ncta) == 0
names = {0:"item start",1:"row start",2:"phasor done",3:"f1+tw done",4:"X1 written",5:"X1 read iss",6:"f2+tw done",7:"X2 done",8:"f3 done",9:"H mul done",10:"i1+tw done",11:"X3 done",12:"i2+tw done",13:"X4 written",14:"X4 read iss",15:"i3 done",16:"G1 tw done/G0 at wait",17:"G1 posted/G0 got mail",18:"G0 emitted",19:"item end"}
# median duration of each phase (slot s-1 -> s) over all CTAs, warps and steady-state items (item >= 1)
print("median phase duration, steady-state rows (item >= 1): group 0 | group 1")
for s_ in range(1, 20):
    row = []
    for grp in (slice(0, 8), slice(8, 16)):
        a_, b_ = buf[:, grp, 1:, s_ - 1], buf[:, grp, 1:, s_]
        ok_ = (a_ > 0) & (b_ > 0)
        row.append(float(np.median((b_ - a_)[ok_])) if ok_.any() else float("nan"))
    print(f"{s_:2d} {names[s_]:24s} {row[0]:8.0f} {row[1]:8.0f}")
# stamps 24..28 sit right after the five token acquires (== the preceding stamp unless built with -DCAF_PINGPONG)
print("token wait (previous stamp -> after acquire) and block length (after acquire -> next stamp): G0 | G1")
for acq, prev_, nxt_ in ((24, 1, 2), (25, 5, 6), (26, 7, 8), (27, 11, 12), (28, 14, 15)):
    row = []
    for grp in (slice(0, 8), slice(8, 16)):
        p_, a_, n_ = buf[:, grp, 1:, prev_], buf[:, grp, 1:, acq], buf[:, grp, 1:, nxt_]
        ok_ = (p_ > 0) & (a_ > 0) & (n_ > 0)
        row += [float(np.median((a_ - p_)[ok_])), float(np.median((n_ - a_)[ok_]))] if ok_.any() else [float("nan")] * 2
    print(f"acquire {acq}: wait {row[0]:6.0f} block {row[1]:6.0f} | wait {row[2]:6.0f} block {row[3]:6.0f}")
for grp, nm in ((slice(0, 8), "G0"), (slice(8, 16), "G1")):
    a_, b_ = buf[:, grp, 1:, 1], buf[:, grp, 1:, 19]
    ok_ = (a_ > 0) & (b_ > 0)
    print(nm, "row start -> item end median", float(np.median((b_ - a_)[ok_])))
a_, b_ = buf[:, 0, 1:-1, 1], buf[:, 0, 2:, 1]
ok_ = (a_ > 0) & (b_ > 0)
print("row start -> next row start (warp 0) median", float(np.median((b_ - a_)[ok_])))
for cta in ((0, 77) if os.environ.get("CAF_TRACE_VERBOSE") else ()):
    t0 = buf[cta][buf[cta] > 0].min()
    print(f"=== CTA {cta}: cycles since first stamp; columns = warp 0 (G0), warp 7 (G0), warp 8 (G1), warp 15 (G1)")
    for item in range(3):
        print(f"-- item {item}")
        prev = None
        for s in range(20):
            vals = [int(buf[cta, w, item, s] - t0) if buf[cta, w, item, s] > 0 else -1 for w in (0, 7, 8, 15)]
            print(f"{s:2d} {names[s]:24s} " + " ".join(f"{v:8d}" for v in vals))
# summary: per-row duration across CTAs
dur = []
for cta in range(ncta):
    for item in range(3):
        a, b = buf[cta, 0, item, 1], buf[cta, 0, item, 19]
        if a > 0 and b > 0: dur.append(b - a)
print("row duration (warp 0) cycles: median", np.median(dur), "min", np.min(dur), "max", np.max(dur))
pro = [buf[c, 0, 0, 1] - buf[c, 0, 0, 0] for c in range(ncta) if buf[c, 0, 0, 1] > 0]
print("prologue (item start -> row start) cycles: median", np.median(pro))

# whole-kernel view: CTA lifetimes on the global timer (ns) and on the SM clock
g0 = buf[:, 0, 0, 20]; g1 = buf[:, 0, 0, 21]; c0 = buf[:, 0, 0, 22]; c1 = buf[:, 0, 0, 23]
ok = g0 > 0
base = g0[ok].min()
print("CTA start (ns after first CTA): min %d med %d max %d" % ((g0[ok]-base).min(), np.median(g0[ok]-base), (g0[ok]-base).max()))
print("CTA end   (ns after first CTA): min %d med %d max %d" % ((g1[ok]-base).min(), np.median(g1[ok]-base), (g1[ok]-base).max()))
life = (c1 - c0)[ok]
print("CTA lifetime cycles: min %d med %d max %d" % (life.min(), np.median(life), life.max()))
first = np.array([buf[c, 0, 0, 0] - buf[c, 0, 0, 22] for c in range(ncta) if buf[c,0,0,0] > 0])
print("setup (entry -> first item start) cycles: med %d max %d" % (np.median(first), first.max()))
nrows = np.array([(buf[c, 0, :, 19] > 0).sum() for c in range(ncta)])
for k in (2, 3):
    sel = (nrows == k) & ok
    if sel.any(): print(f"CTAs with {k} rows: n={sel.sum()} lifetime med {np.median((c1-c0)[sel]):.0f} max {(c1-c0)[sel].max()}")

# token intervals of one CTA in absolute SM cycles (exclusive if the ping-pong token works): item 1, warps 0 / 8
if os.environ.get("CAF_TRACE_TOKEN"):
    cta = 77
    ev = []
    for wp, nm, last in ((0, "G0", 18), (8, "G1", 16)):
        for it in (1, 2):
            for acq, rel in ((24, 3), (25, 6), (26, 10), (27, 12), (28, last)):
                if buf[cta, wp, it, acq] > 0: ev.append((int(buf[cta, wp, it, acq]), int(buf[cta, wp, it, rel]), nm, it, acq))
    ev.sort()
    t0 = ev[0][0]
    for a_, r_, nm, it, acq in ev: print(f"{nm} item {it} acquire {acq}: [{a_ - t0:7d}, {r_ - t0:7d})  len {r_ - a_}")

# how many of the 16 warps are inside a butterfly block at any time (steady-state rows): 0 = fp64 pipe idle
if os.environ.get("CAF_TRACE_OCC"):
    comp = ((1, 2), (2, 3), (5, 6), (7, 8), (8, 9), (9, 10), (11, 12), (14, 15), (15, 16), (17, 18))
    hist = np.zeros(17)
    for cta in range(ncta):
        a0, a1 = buf[cta, 0, 1, 1], buf[cta, 0, 2, 1]       # row start of item 1 .. row start of item 2 (warp 0)
        if a0 <= 0 or a1 <= 0: continue
        n = int(a1 - a0)
        occ = np.zeros(n + 1, dtype=np.int32)
        for wp in range(16):
            for it in (0, 1, 2):
                for s0, s1 in comp:
                    b0, b1 = buf[cta, wp, it, s0], buf[cta, wp, it, s1]
                    if b0 <= 0 or b1 <= 0: continue
                    lo_, hi_ = int(max(b0 - a0, 0)), int(min(b1 - a0, n))
                    if hi_ > lo_:
                        occ[lo_] += 1; occ[hi_] -= 1
        occ = np.cumsum(occ)[:n]
        hist += np.bincount(occ, minlength=17)[:17]
    hist /= hist.sum()
    print("fraction of row time with k warps inside a butterfly block:")
    print("  k=0 %.3f | 1-4 %.3f | 5-8 %.3f | 9-12 %.3f | 13-16 %.3f" % (hist[0], hist[1:5].sum(), hist[5:9].sum(), hist[9:13].sum(), hist[13:].sum()))
    print("  mean warps computing %.2f" % (np.arange(17) * hist).sum())

# H producers (CTA 0 group 0, CTA hprod1 group 1): the haystack transform is stamped into item slot 7
if os.environ.get("CAF_TRACE_H"):
    for cta in range(ncta):
        for wp in (0, 8):
            if buf[cta, wp, 7, 0] > 0:
                c0_ = buf[cta, 0, 0, 22]
                st = [int(buf[cta, wp, 7, s_] - c0_) if buf[cta, wp, 7, s_] > 0 else -1 for s_ in (0, 3, 4, 5, 6, 7, 8, 9)]
                print(f"H producer CTA {cta} warp {wp}: cycles since CTA entry: start {st[0]} f1 {st[1]} X1w {st[2]} X1r {st[3]} f2 {st[4]} X2 {st[5]} f3 {st[6]} published {st[7]}")
    # consumers: when did H arrive (slot 9 of item 0) relative to CTA entry
    arr = [int(buf[c_, 0, 0, 9] - buf[c_, 0, 0, 22]) for c_ in range(ncta) if buf[c_, 0, 0, 9] > 0 and buf[c_, 0, 7, 0] == 0]
    f3 = [int(buf[c_, 0, 0, 8] - buf[c_, 0, 0, 22]) for c_ in range(ncta) if buf[c_, 0, 0, 9] > 0 and buf[c_, 0, 7, 0] == 0]
    print("consumers (G0): f3 done at median %d, H multiplied at median %d cycles after CTA entry" % (np.median(f3), np.median(arr)))
