import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")
FS = 48000


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests fail loudly on a box without a GPU instead of silently skipping or falling back."""
    # nothing to rewrite: GPU tests create a Handle, which raises CafError(ENODEVICE) when no sm_100 device exists


@pytest.fixture(scope="session")
def known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)["cases"]


def load_case(case):
    """(needle, haystack[:len(needle)], shifts) exactly as caf_rust/tests/test.rs prepares them."""
    from oracle import oracle as O
    needle = O.read_file_c64(os.path.join(DATA, "chirp_%d_raw.c64" % case["chirp"]))
    hay = O.read_file_c64(os.path.join(DATA, case["haystack"]))[: needle.size]
    shifts = O.gen_float_shifts(*case["grid"])
    return needle, hay, shifts


@pytest.fixture(scope="session")
def chirp0():
    from oracle import oracle as O
    needle = O.read_file_c64(os.path.join(DATA, "chirp_0_raw.c64"))
    hay = O.read_file_c64(os.path.join(DATA, "chirp_0_T+202samp_F+69.25Hz.c64"))[: needle.size]
    return needle, hay


def rel_max(a, b):
    """The tolerance metric of this repo: max|a-b| / max|b|."""
    a = np.asarray(a); b = np.asarray(b)
    den = float(np.abs(b).max())
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())
