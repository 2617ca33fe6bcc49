"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/caf_b200.h declares,
fails loudly without a GPU, and the pure-host helpers behave.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from caf_cookoff_b200 import _lib, api
from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "caf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(caf_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.SO_PATH), "run __graft_entry__.build()"
    assert os.path.dirname(_lib.SO_PATH).endswith("caf_cookoff_b200")


def test_every_declared_symbol_is_exported_and_typed():
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/caf_b200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in _lib.SYMBOLS"
    assert sorted(_lib.SYMBOLS) == declared


def test_sm100a_code_only():
    """The library carries sm_100a SASS and nothing else (no multi-arch fallback)."""
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_peak_struct_layout():
    assert C.sizeof(_lib.Peak) == 32
    assert [f[0] for f in _lib.Peak._fields_] == ["value", "freq_hz", "doppler_idx", "delay_idx"]


def test_version_and_error_string():
    lib = _lib.load()
    assert b"sm_100a" in lib.caf_b200_version()
    assert isinstance(lib.caf_b200_last_error(), bytes)


def test_no_gpu_means_loud_failure_not_fallback():
    """Without an sm_100 device handle creation must fail with ENODEVICE — there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU-only box")
    with pytest.raises(api.CafError) as e:
        api.Handle(0)
    assert e.value.status == -5
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(api.CafError):
        api.CafB200.caf_surface(np.zeros(8, complex), np.zeros(8, complex), [0.0], 48000)


def test_null_handle_is_rejected():
    lib = _lib.load()
    assert lib.caf_b200_sync(None) == -1
    assert lib.caf_b200_destroy(None) == 0
    assert lib.caf_b200_launch_count(None) == 0


def test_peak_pack_resolve_is_find_peak_over_shards():
    """Sharded find_peak (mod.rs:31-42): the global winner is the largest value, ties to the lowest global row."""
    def pk(v, f, d, k):
        return _lib.Peak(v, f, d, k)
    none = _lib.Peak(0.0, 0.0, api.UINT64_MAX, 0)
    words = np.stack([api.peak_pack(pk(5.0, 1.5, 3, 77), 0), api.peak_pack(pk(9.0, 2.5, 1, 11), 100),
                      api.peak_pack(pk(9.0, 3.5, 0, 12), 200), api.peak_pack(none, 300)])
    out = api.peak_resolve(words)
    assert (out.value, out.freq_hz, out.doppler_idx, out.delay_idx) == (9.0, 2.5, 101, 11)
    out = api.peak_resolve(np.stack([api.peak_pack(none, 0), api.peak_pack(none, 50)]))
    assert (out.value, out.freq_hz, out.doppler_idx, out.delay_idx) == (0.0, 0.0, api.UINT64_MAX, 0)
    # order of the shards must not matter
    out2 = api.peak_resolve(words[::-1].copy())
    assert (out2.doppler_idx, out2.delay_idx) == (101, 11)


def test_nccl_is_bound_at_run_time_and_id_works_without_a_gpu():
    """caf_b200_comm_*: libnccl is dlopen()ed lazily (no link-time dependency) and an id can be minted on the CPU."""
    import shutil, subprocess
    lib = _lib.load()
    buf = (C.c_ubyte * 128)()
    rc = lib.caf_b200_comm_unique_id(C.cast(buf, C.c_void_p))
    assert rc == 0, lib.caf_b200_last_error()
    assert any(bytes(buf))
    out = C.c_void_p()
    assert lib.caf_b200_comm_create(None, 2, 0, C.cast(buf, C.c_void_p), C.byref(out)) == -1     # EINVAL: null handle
    ldd = shutil.which("ldd")
    if ldd:
        deps = subprocess.run([ldd, _lib.SO_PATH], capture_output=True, text=True).stdout
        assert "nccl" not in deps and "torch" not in deps, deps
