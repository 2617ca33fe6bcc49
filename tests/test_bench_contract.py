"""bench.py's reference arm is the one leg of the bench contract that runs without a GPU: it times the oracle
port of CafRustFFTThreadpool (caf_rust/src/caf/mod.rs:391-461) on the host cores.  These tests hold its JSON line
to the contract (one line on stdout, `impl: reference`, the b200 arm's metric / unit / config keys, an `e2e`
with no copies, a `cpu_baseline` describing the run) and check that only rank 0 works under torchrun."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline")


def run_bench(extra_env=None, *args):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = run_bench(None, "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in REQUIRED:
        assert key in line, key
    assert line["unit"] == "cells/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["dtype"] == "f64" and "workload" in line["config"] and "model" not in line["config"]
    # 2 surfaces of 400 x 8192 cells in ms_per_step each
    assert abs(line["value"] - 400 * 8192 / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    r = run_bench({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""
