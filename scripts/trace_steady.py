"""Development: phase timeline of the row kernel in STEADY STATE (a batch launch: whole pairs per CTA, rows 4..6 of a CTA),
all 16 warps of one CTA side by side.  Needs the -DCAF_TRACE build:
   CAF_B200_SO=devlibs/lib_trace.so python scripts/trace_steady.py"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from caf_cookoff_b200 import _lib, read_file_c64, bench_shifts, Handle
D = os.path.join(ROOT, "tests/golden/data/")
needle = read_file_c64(D + "chirp_0_raw.c64"); hay = read_file_c64(D + "chirp_0_T+202samp_F+69.25Hz.c64")[:4096]
sh = bench_shifts()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
lib.caf_b200_debug_trace(h.raw, None, 0)            # allocate
P = 148
nd = torch.from_numpy(needle).to(dev).repeat(P, 1).contiguous(); hd = torch.from_numpy(hay).to(dev).repeat(P, 1).contiguous()
fd = torch.from_numpy(sh).to(dev); pk = torch.zeros(4 * P, dtype=torch.int64, device=dev)
for _ in range(2):
    rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), P, 4096, fd.data_ptr(), sh.size, 48000, None, None, None, pk.data_ptr())
    assert rc == 0
torch.cuda.synchronize()
ncta = 148
buf = np.zeros((ncta, 16, 8, 32), dtype=np.int64)
assert lib.caf_b200_debug_trace(h.raw, buf.ctypes.data_as(C.c_void_p), ncta) == 0
names = {0:"item start",1:"row start",2:"phasor/f1",3:"f1/tw done",4:"X1 written",5:"X1 read",6:"f2 done",7:"X2 done",8:"f3 done",9:"H done",10:"i1 done",11:"X3 done",12:"i2 done",13:"X4 written",14:"X4 read",15:"i3 done",16:"G1 tw/G0 wait",17:"G1 post/G0 mail",18:"G0 emitted",19:"item end"}
cta = int(os.environ.get("CTA", "77"))
t0 = buf[cta, :, 4, 1].min()
print(f"CTA {cta}: cycles since row 4 start; columns: G0 warps 0..7 | G1 warps 8..15")
for item in (4, 5):
    print(f"-- item {item}")
    for s in range(1, 20):
        vals = [int(buf[cta, w, item, s] - t0) if buf[cta, w, item, s] > 0 else -1 for w in range(16)]
        print(f"{s:2d} {names[s]:16s} " + " ".join(f"{v:6d}" for v in vals[:8]) + " | " + " ".join(f"{v:6d}" for v in vals[8:]))
# row period per group (row start to row start), medians over CTAs
for g, nm in ((0, "G0"), (8, "G1")):
    per = buf[:, g, 5:8, 1] - buf[:, g, 4:7, 1]
    print(nm, "row period median", float(np.median(per)))
# compute-occupancy histogram: number of warps inside a compute block over rows 4..6
comp = ((1, 2), (2, 3), (5, 6), (7, 8), (8, 9), (9, 10), (11, 12), (14, 15), (15, 16), (17, 18))
hist = np.zeros(17)
for c_ in range(ncta):
    a0, a1 = buf[c_, 0, 4, 1], buf[c_, 0, 7, 1]
    if a0 <= 0 or a1 <= 0: continue
    n = int(a1 - a0)
    occ = np.zeros(n + 1, dtype=np.int32)
    for wp in range(16):
        for it in range(3, 8):
            for s0, s1 in comp:
                b0, b1 = buf[c_, wp, it, s0], buf[c_, wp, it, s1]
                if b0 <= 0 or b1 <= 0: continue
                lo_, hi_ = int(max(b0 - a0, 0)), int(min(b1 - a0, n))
                if hi_ > lo_:
                    occ[lo_] += 1; occ[hi_] -= 1
    occ = np.cumsum(occ)[:n]
    hist += np.bincount(occ, minlength=17)[:17]
hist /= hist.sum()
print("fraction of time with k warps inside a compute block: " + " ".join(f"{k}:{hist[k]:.3f}" for k in range(17)))
print("mean warps computing %.2f" % (np.arange(17) * hist).sum())
