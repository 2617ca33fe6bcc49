#!/usr/bin/env python
"""bench.py — the reference's headline benchmark (caf_rust/benches/caf_bench.rs:150-168: one 400 x 8192
fp64 CAF surface + find_peak on the chirp_0 pair) on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg1|cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one surface: FFT(s1) once, the fused row kernel over all doppler
rows, find_peak.  At N > 1 every rank owns its own s0/s1 pair (pairs sharded, no data-path collective:
"scaling": "weak").  `value` is device-resident throughput (inputs already in HBM), `e2e` goes through the
host-pointer C ABI (pinned host buffers, H2D + kernels + D2H of the whole surface every step).
Between timed steps L2 is flushed by overwriting a 256 MiB buffer; each step is timed with its own pair
of CUDA events on the launching stream and the K durations are summed (max over ranks).
At N = 1 the line also carries `working_set_gt_l2`: the same kernel over a working set larger than L2 (320 seeded
pairs, 8 surface buffers) with one event pair around the back-to-back launches (scripts/bench_stream.py) —
informational, `value` and `roofline` stay on the flushed per-step figure.

--impl reference times the reference's CPU algorithm (oracle port of CafRustFFTThreadpool, all host
cores) on the same workload; the Rust crate itself cannot be compiled in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
L = 4096
N = 2 * L
# the reference's published figure for exactly this metric (one 400 x 8192 fp64 surface +
# find_peak, README.md:36, rust RustFFT + threadpool on an R9-3900X, 12C/24T): 28 ms = 117.0 Mcell/s.  Other hardware.
PUBLISHED_CELLS_PER_S = 400 * N / 28e-3
PUBLISHED_NOTE = "BASELINE.md: rust RustFFT + threadpool, 28 ms per surface on an R9-3900X (README.md:36)"


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner on the first
    communicator), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def algorithmic_flops(d: int, n: int) -> float:
    """per row 10 N log2 N + 15 N; plus 5 N log2 N once per pair for FFT(s1)."""
    lg = np.log2(n)
    return d * (10.0 * n * lg + 15.0 * n) + 5.0 * n * lg


def load_pair(index: int):
    """s0/s1 of the reference benchmark (chirp_0) or, for other ranks, the next pairs of the same seed-0
    stream — read from the committed fixtures, which the seeded generator port reproduces bit-exactly."""
    from caf_cookoff_b200 import read_file_c64
    data = os.path.join(ROOT, "tests", "golden", "data")
    index %= 10
    names = sorted(f for f in os.listdir(data) if f.startswith(f"chirp_{index}_T"))
    needle = read_file_c64(os.path.join(data, f"chirp_{index}_raw.c64"))
    hay = read_file_c64(os.path.join(data, names[0]))
    hay = np.resize(hay[: needle.size], needle.size) if hay.size >= needle.size else np.concatenate(
        [hay, np.zeros(needle.size - hay.size, dtype=hay.dtype)])   # caf_bench.rs:28 haystack.resize
    return needle, hay


class ClockSampler:
    """nvidia-smi style clock / throttle-reason sampling during the timed region (NVML)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report that, do not invent clocks
            self.nv, self.err = None, repr(e)

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def sample(self):
        if not self.nv:
            return
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM)))
            r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
            for bit, name in self._NAMES.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(0.002)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml")}
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_cpus(index: int):
    """One process per GPU: run (and first-touch the pinned host buffers) on the CPUs NVML reports as local to this
    GPU, so the 26 MB D2H copy of every e2e step does not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def committed_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the row kernel, per launch, from the newest committed
    `ncu --set full` capture (profiles/rNN_traffic.json, written by scripts/summarize_ncu.py)."""
    try:
        pdir = os.path.join(ROOT, "profiles")
        names = sorted(n for n in os.listdir(pdir) if n.endswith("_traffic.json"))
        return json.load(open(os.path.join(pdir, names[-1]))).get("dram_bytes_per_launch") if names else None
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def cpu_baseline_run(needle, hay, freqs, budget_s: float, threads: int):
    """The oracle port of CafRustFFTThreadpool (mod.rs:391-461) on this box's host cores."""
    from oracle import oracle as O
    O.caf_surface(needle, hay, freqs[:8], FS, threads=threads, want_surface=True)   # warm (page-in, plan)
    times = []
    t_end = time.perf_counter() + budget_s
    while True:
        t0 = time.perf_counter()
        surf, pidx, pval = O.caf_surface(needle, hay, freqs, FS, threads=threads, want_surface=True)
        O.find_peak(freqs, pidx, pval)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() >= t_end and len(times) >= 3:
            break
    return float(np.median(times)), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    needle, hay = load_pair(0)
    from caf_cookoff_b200 import bench_shifts
    freqs = bench_shifts()
    from oracle import oracle as O
    for _ in range(max(args.warmup, 1)):
        O.caf_surface(needle, hay, freqs, FS, threads=threads, want_surface=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        surf, pidx, pval = O.caf_surface(needle, hay, freqs, FS, threads=threads, want_surface=True)
        O.find_peak(freqs, pidx, pval)
    dt = time.perf_counter() - t0
    cells = freqs.size * N
    value = cells * args.steps / dt
    line = {
        "impl": "reference", "metric": "CAF cells/s (400x8192 fp64 surface + peak)", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": value / PUBLISHED_CELLS_PER_S,
        "vs_baseline_note": PUBLISHED_NOTE, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg1: 400 doppler x 8192 delay fp64 CAF surface + peak, chirp_0 pair (seed 0), fs=48000"},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} whole surfaces, oracle port of CafRustFFTThreadpool (3 FFTs/row), {threads} threads"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from caf_cookoff_b200 import Handle, _lib, bench_shifts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: torch.cuda.is_available() is False and there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_cpus(local)       # pinned staging buffers then live on the socket the GPU hangs off
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    f32 = args.workload == "cfg2"
    cdt, rdt = (np.complex64, np.float32) if f32 else (np.complex128, np.float64)
    tcdt, trdt = (torch.complex64, torch.float32) if f32 else (torch.complex128, torch.float64)
    sfx = "f32" if f32 else "f64"

    needle, hay = load_pair(rank)
    freqs = bench_shifts()
    D = freqs.size
    stream = torch.cuda.Stream(device=dev)     # every kernel, copy and event of this bench lives on this stream
    torch.cuda.set_stream(stream)
    h = Handle(local, stream=stream.cuda_stream)

    # ---- device-resident inputs / outputs -----------------------------------------------------------
    needle_d = torch.from_numpy(needle.astype(cdt)).to(dev)
    hay_d = torch.from_numpy(hay.astype(cdt)).to(dev)
    freqs_d = torch.from_numpy(freqs).to(dev)
    surf_d = torch.empty((D, N), dtype=trdt, device=dev)
    rv_d = torch.empty(D, dtype=trdt, device=dev)
    ri_d = torch.empty(D, dtype=torch.int64, device=dev)
    pk_d = torch.zeros(4, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 2 x the 126 MB L2
    dev_fn = getattr(lib, f"caf_b200_batch_{sfx}_dev")

    def step_dev():
        rc = dev_fn(h.raw, needle_d.data_ptr(), hay_d.data_ptr(), 1, L, freqs_d.data_ptr(), D, FS,
                    surf_d.data_ptr(), rv_d.data_ptr(), ri_d.data_ptr(), pk_d.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.caf_b200_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_pass(step, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for s0, s1 in evs:
            flush.zero_()
            s0.record(stream)
            step()
            s1.record(stream)
        return evs

    # ---- warm-up, then the timed region ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    launches0 = h.launch_count
    sampler.start()
    barrier()
    evs = timed_pass(step_dev, args.steps)
    sampler.sample()
    barrier()
    sampler.stop()
    launches = h.launch_count - launches0
    per_step = np.array([a.elapsed_time(b) for a, b in evs])   # ms
    total_ms = float(per_step.sum())
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    cells_step = D * N
    value = world * cells_step * args.steps / (total_ms_max * 1e-3)

    # correctness of what was just timed: the peak must be the known answer of this pair
    pk = pk_d.cpu().numpy()
    peak_freq = float(pk.view(np.float64)[1]); peak_delay = int(pk.view(np.uint64)[3])

    # ---- roofline pass: the row kernel alone, same flush regimen, CUDA events inside the library ----------
    lib.caf_b200_set_profiling(h.raw, 1)
    rows_ms, spec_ms, peak_ms = [], [], []
    a_, b_, c_ = C.c_float(), C.c_float(), C.c_float()
    for _ in range(min(max(args.steps, 10), 200)):
        flush.zero_()
        step_dev()
        lib.caf_b200_last_kernel_ms(h.raw, C.byref(a_), C.byref(b_), C.byref(c_))
        spec_ms.append(a_.value); rows_ms.append(b_.value); peak_ms.append(c_.value)
    lib.caf_b200_set_profiling(h.raw, 0)
    rows_avg_ms = float(np.mean(rows_ms))
    tf = C.c_double()
    lib.caf_b200_probe_fma_tflops(h.raw, 0 if f32 else 1, C.byref(tf))
    row_flops = D * (10.0 * N * np.log2(N) + 15.0 * N)
    achieved_tf = row_flops / (rows_avg_ms * 1e-3) / 1e12
    mp = measured_peaks()
    hbm_peak = (mp or {}).get("hbm_gbs", 6650.0)
    surf_bytes = D * N * np.dtype(rdt).itemsize
    roofline = {
        "bound": "fp64" if not f32 else "fp32", "kernel": f"caf_rows_kernel<{'float' if f32 else 'double'}, kSurface>",
        "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved_tf / tf.value if tf.value else None,
        "peak_source": "in-run FMA-pipe probe (caf_b200_probe_fma_tflops); MEASURED_PEAKS.json has no fp64/fp32 entry",
        "flops_per_launch": row_flops, "kernel_ms": rows_avg_ms,
        "traffic": committed_traffic() if not f32 else None,
        "traffic_note": "DRAM bytes per launch from the committed ncu capture; the 26.2 MB surface is written back from the 126 MB L2 after the kernel ends",
        "hbm": {"achieved": surf_bytes / (rows_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": surf_bytes / (rows_avg_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json (measured)" if mp else "fallback 6650 GB/s"},
        "step_share": {"spectrum_ms": float(np.mean(spec_ms)), "rows_ms": rows_avg_ms, "peak_ms": float(np.mean(peak_ms))},
    }
    # the other roofline the north star names: shared memory.  A row moves every value of its two pipelines through the
    # exchange fabric 4 times each way plus the radix-2 mailbox: 72 stores and 72 loads of one complex value per thread
    # (DESIGN.md section 3); nominal peak 128 B/clk per SM at the SM clock.
    try:
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_hz = 1e6 * float(sampler.max_mhz or (mp or {}).get("sm_max_mhz", 1965.0))
        smem_bytes_row = 512 * 144 * (8 if f32 else 16)
        smem_achieved = D * smem_bytes_row / (rows_avg_ms * 1e-3) / 1e12
        smem_peak = sm_count * 128.0 * sm_hz / 1e12
        roofline["smem"] = {"achieved": smem_achieved, "peak": smem_peak, "unit": "TB/s", "frac": smem_achieved / smem_peak,
                            "bytes_per_row": smem_bytes_row,
                            "note": "nominal 128 B/clk/SM; measured LDS.128 runs at 64 B/clk, STS.128 at 112 (scripts/micro/mio_cost.cu)"}
    except Exception as e:      # never let the extra figure break the bench line
        roofline["smem"] = {"error": repr(e)}

    # ---- e2e: the host-pointer C ABI with pinned host buffers, H2D + D2H inside the timed region ----------
    needle_h = torch.from_numpy(needle.astype(cdt)).pin_memory()
    hay_h = torch.from_numpy(hay.astype(cdt)).pin_memory()
    freqs_h = torch.from_numpy(freqs).pin_memory()
    surf_h = torch.empty((D, N), dtype=trdt).pin_memory()
    rv_h = torch.empty(D, dtype=trdt).pin_memory()
    ri_h = torch.empty(D, dtype=torch.int64).pin_memory()
    pk_h = _lib.Peak()
    host_fn = getattr(lib, f"caf_b200_surface_{sfx}")

    def step_host():
        rc = host_fn(h.raw, needle_h.data_ptr(), hay_h.data_ptr(), L, freqs_h.data_ptr(), D, FS,
                     surf_h.data_ptr(), rv_h.data_ptr(), ri_h.data_ptr(), C.cast(C.byref(pk_h), C.c_void_p))
        if rc != 0:
            raise RuntimeError(lib.caf_b200_last_error().decode())

    e2e_steps = min(args.steps, 200)
    for _ in range(3):
        step_host()
    barrier()
    t0 = time.perf_counter()
    evs = timed_pass(step_host, e2e_steps)
    barrier()
    wall = time.perf_counter() - t0
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(t.item())
    h2d = needle_h.numel() * needle_h.element_size() + hay_h.numel() * hay_h.element_size() + freqs_h.numel() * 8
    d2h = surf_h.numel() * surf_h.element_size() + rv_h.numel() * rv_h.element_size() + ri_h.numel() * 8 + 32
    e2e = {"value": world * cells_step * e2e_steps / (e2e_ms_max * 1e-3), "unit": "cells/s",
           "ms_per_step": e2e_ms_max / e2e_steps, "steps": e2e_steps, "wall_ms_per_step_incl_flush": 1e3 * wall / e2e_steps,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "peak": [pk_h.freq_hz, int(pk_h.delay_idx)]}

    # ---- the same call without the surface crossing PCIe: caf_b200_peak_* (what caf_bench.rs's closure observes:
    #      find_peak(caf_surface(..)) returns (freq, delay); CafSurfaceRow's fields are private, mod.rs:17-22) --------
    peak_fn = getattr(lib, f"caf_b200_peak_{sfx}")
    pk2 = _lib.Peak()

    def step_peak():
        rc = peak_fn(h.raw, needle_h.data_ptr(), hay_h.data_ptr(), L, freqs_h.data_ptr(), D, FS,
                     C.cast(C.byref(pk2), C.c_void_p))
        if rc != 0:
            raise RuntimeError(lib.caf_b200_last_error().decode())

    for _ in range(3):
        step_peak()
    barrier()
    evs = timed_pass(step_peak, e2e_steps)
    barrier()
    pk_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    t = torch.tensor([pk_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_peak = {"value": world * cells_step * e2e_steps / (float(t.item()) * 1e-3), "unit": "cells/s",
                "ms_per_step": float(t.item()) / e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 32,
                "peak": [pk2.freq_hz, int(pk2.delay_idx)],
                "note": "host inputs in, (freq, delay) out: the surface stays on the GPU (caf_b200_peak_*)"}

    # ---- second timing method (N = 1 only; last thing that touches the GPU): working set larger than L2, K back-to-back
    #      launches in ONE event pair, so the ~6 us every kernel pays between two events is not charged to each step
    #      (scripts/bench_stream.py; measured there at 43.1 us per surface against 47.7 us with per-step event pairs).
    #      Informational: `value` and `roofline` above stay on the flushed per-step figure. ---------------------------
    stream_fig = None
    if rank == 0 and world == 1:
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import bench_stream
            stream_fig = bench_stream.measure(lib, h, stream, dev, pairs=320, surfaces=8, steps=min(max(args.steps, 50), 300),
                                              warmup=40, f32=f32)
        except BaseException as e:      # never let the extra figure break the bench line
            stream_fig = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core again
        cores = len(all_cpus) or 1
        med, reps = cpu_baseline_run(needle, hay, freqs, args.cpu_budget_s, cores)
        cpu = {"value": cells_step / med, "unit": "cells/s", "ms_per_surface": med * 1e3, "cores": cores, "kind": "port",
               "sample": f"{reps} whole 400x8192 surfaces (median), oracle port of CafRustFFTThreadpool, {cores} threads"}

    if rank == 0:
        line = {
            "metric": "CAF cells/s (400x8192 %s surface + peak)" % ("fp32" if f32 else "fp64"),
            "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms_max / args.steps, "ms_per_surface": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None if f32 else value / PUBLISHED_CELLS_PER_S,
            "vs_baseline_note": "no published complex64 figure" if f32 else PUBLISHED_NOTE,
            "dtype": "f32" if f32 else "f64", "data": "synthetic",
            "config": {"workload": ("cfg2" if f32 else "cfg1") + ": 400 doppler x 8192 delay CAF surface + peak per step per GPU, "
                       "utils/generate.py seed-0 pairs (rank r uses chirp_r), fs=48000",
                       "doppler_rows": D, "delay_cells": N, "pairs_per_step_per_gpu": 1,
                       "l2": "flushed between timed steps (256 MiB overwrite); each step timed with its own CUDA event pair",
                       "parallelism": f"pairs sharded x{world}, no data-path collective",
                       "host_cpus_bound_to_gpu": numa},
            "e2e": e2e, "e2e_peak_only": e2e_peak, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "working_set_gt_l2": stream_fig,
            "clocks": sampler.summary(),
            "check": {"peak_freq_hz": peak_freq, "peak_delay": peak_delay},
            "step_ms_min_med_max": [float(per_step.min()), float(np.median(per_step)), float(per_step.max())],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=["cfg1", "cfg2"], default="cfg1")
    ap.add_argument("--cpu-budget-s", type=float, default=3.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
