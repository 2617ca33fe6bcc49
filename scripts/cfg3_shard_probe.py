"""Development: one rank's share of config 3 at N = 8 (512 doppler rows of 65536 cells) on one GPU through the sharded
entry (world-1 communicator): step time by CUDA events, to compare with the sum of the kernels' durations under ncu."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from caf_cookoff_b200 import Handle, _lib, generate as G, dist as cdist
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
lib = _lib.load(); h = Handle(0, stream=stream.cuda_stream)
L = 32768
needle, hay = G.as_inputs(G.pair(0, seed=0, chirp_length=L))
freqs = np.linspace(-100.0, 100.0, 4096, endpoint=False)[:rows].copy()
comm = cdist.Comm(h, 1, 0, os.path.join(tempfile.gettempdir(), "caf_probe_id_%d" % os.getpid()))
nd = torch.from_numpy(needle).to(dev); hd = torch.from_numpy(hay).to(dev); fd = torch.from_numpy(freqs).to(dev)
surf = torch.empty((rows, 2 * L), dtype=torch.float64, device=dev)
rv = torch.empty(rows, dtype=torch.float64, device=dev); ri = torch.empty(rows, dtype=torch.int64, device=dev)
outp = torch.zeros(4, dtype=torch.int64, device=dev)
def step():
    comm.sharded_dev(nd.data_ptr(), hd.data_ptr(), L, fd.data_ptr(), rows, 0, 48000, outp.data_ptr(),
                     surface_local_dev=surf.data_ptr(), row_val_dev=rv.data_ptr(), row_idx_dev=ri.data_ptr())
for _ in range(3): step()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps): step()
e1.record(stream); torch.cuda.synchronize()
print(f"rows {rows}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us per step ({steps} steps), launches per step {h.launch_count // (steps + 3)}")
comm.close()
