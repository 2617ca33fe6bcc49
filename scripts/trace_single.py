"""Development: where the time of ONE 400-row surface launch goes (needs the -DCAF_TRACE build).
   CAF_B200_SO=devlibs/lib_trace.so python scripts/trace_single.py"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from caf_cookoff_b200 import _lib, read_file_c64, bench_shifts, Handle
D_ = os.path.join(ROOT, "tests/golden/data/")
needle = read_file_c64(D_ + "chirp_0_raw.c64"); hay = read_file_c64(D_ + "chirp_0_T+202samp_F+69.25Hz.c64")[:4096]
sh = bench_shifts(); D = sh.size
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
lib.caf_b200_debug_trace(h.raw, None, 0)
nd = torch.from_numpy(needle).to(dev); hd = torch.from_numpy(hay).to(dev); fd = torch.from_numpy(sh).to(dev)
surf = torch.empty((D, 8192), dtype=torch.float64, device=dev); rv = torch.empty(D, dtype=torch.float64, device=dev)
ri = torch.empty(D, dtype=torch.int64, device=dev); pk = torch.zeros(4, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ncta = 148
buf = np.zeros((ncta, 16, 8, 32), dtype=np.int64)
res = []
for i in range(12):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, 4096, fd.data_ptr(), D, 48000, surf.data_ptr(), rv.data_ptr(), ri.data_ptr(), pk.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    assert rc == 0
    if i < 4: continue
    assert lib.caf_b200_debug_trace(h.raw, buf.ctypes.data_as(C.c_void_p), ncta) == 0
    g0 = buf[:, 0, 0, 20]; g1 = buf[:, 0, 0, 21]; g2 = buf[:, 0, 0, 29]; c0 = buf[:, 0, 0, 22]; c1 = buf[:, 0, 0, 23]
    base = g0.min()
    ghz = np.median((c1 - c0) / np.maximum(g1 - g0, 1))
    nrows = np.array([(buf[c_, 0, :, 19] > 0).sum() for c_ in range(ncta)])
    three = nrows == 3
    def cyc(a_): return float(np.median(a_))
    # group 0 warp 0 and group 1 warp 8 stamps, CTAs with three rows
    w0 = buf[three, 0]; w8 = buf[three, 8]
    ent = c0[three]
    line = {
        "events_us": e0.elapsed_time(e1) * 1e3,
        "first_entry_to_last_exit_us": (g2.max() - base) / 1e3, "last_row_done_us": (g1.max() - base) / 1e3,
        "ghz": ghz,
        "setup": cyc(w0[:, 0, 0] - ent), "prologue": cyc(w0[:, 0, 1] - w0[:, 0, 0]),
        "row1_G0": cyc(w0[:, 0, 19] - w0[:, 0, 1]), "row2_G0": cyc(w0[:, 1, 19] - w0[:, 1, 1]), "row3_G0": cyc(w0[:, 2, 19] - w0[:, 2, 1]),
        "row1_wait_H": cyc(w0[:, 0, 9] - w0[:, 0, 8]),
        "gap12": cyc(w0[:, 1, 1] - w0[:, 0, 19]), "gap23": cyc(w0[:, 2, 1] - w0[:, 1, 19]),
        "G0_last_row_end": cyc(w0[:, 2, 19] - ent), "G1_last_row_end": cyc(w8[:, 2, 19] - ent),
        "lifetime3": cyc((c1 - c0)[three]), "lifetime2": cyc((c1 - c0)[nrows == 2]) if (nrows == 2).any() else -1,
    }
    res.append(line)
keys = res[0].keys()
print({k: round(float(np.median([r_[k] for r_ in res])), 2) for k in keys})
