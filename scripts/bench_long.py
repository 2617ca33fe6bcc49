"""Development timing of the long-row path (BASELINE config 3: 4096 doppler x 65536 delay, fp64), device resident."""
import os, sys, time, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from caf_cookoff_b200 import Handle, _lib, generate as G
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
want_surface = (len(sys.argv) <= 3) or sys.argv[3] != "peak"
dev = torch.device("cuda", 0); stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h = Handle(0, stream=stream.cuda_stream); lib = _lib.load()
needle, hay = G.as_inputs(G.pair(0, chirp_length=L))
freqs = np.linspace(-100, 100, D, endpoint=False)
nd = torch.from_numpy(needle).to(dev); hd = torch.from_numpy(hay).to(dev); fd = torch.from_numpy(freqs).to(dev)
surf = torch.empty((D, 2 * L), dtype=torch.float64, device=dev) if want_surface else None
rv = torch.empty(D, dtype=torch.float64, device=dev); ri = torch.empty(D, dtype=torch.int64, device=dev); pk = torch.zeros(4, dtype=torch.int64, device=dev)
def step():
    rc = lib.caf_b200_batch_f64_dev(h.raw, nd.data_ptr(), hd.data_ptr(), 1, L, fd.data_ptr(), D, 48000,
                                    surf.data_ptr() if want_surface else None, rv.data_ptr(), ri.data_ptr(), pk.data_ptr())
    assert rc == 0, lib.caf_b200_last_error()
for _ in range(2): step()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
n = 3
e0.record(stream)
for _ in range(n): step()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
N = 2 * L; flops = D * (10.0 * N * np.log2(N) + 15.0 * N)
p = pk.cpu().numpy()
print(f"L={L} D={D} surface={want_surface}: {ms:.3f} ms  {D*N/ms/1e6:.1f} Gcell/s  {flops/ms/1e9:.2f} TFLOP/s (algorithmic)  surface write {D*N*8/ms/1e6 if want_surface else 0:.0f} GB/s  peak f={p.view(np.float64)[1]} delay={p.view(np.uint64)[3]}")
