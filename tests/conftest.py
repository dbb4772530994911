"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/): test infrastructure only."""
    from oracle import oracle_py
    oracle_py.build(ref=None if not oracle_py.have_ref() else False)
    return oracle_py


@pytest.fixture(scope="session")
def gpu_ctx():
    from pywfa_b200 import _ffi
    from pywfa_b200.build import build_library
    build_library()
    return _ffi.Context(0)


def assert_same(got, want, scope_full=True, what=""):
    """Bit-exact comparison of score, status, CIGAR runs and coordinates."""
    for key in ("score", "status"):
        bad = np.flatnonzero(got[key] != want[key])
        assert bad.size == 0, f"{what}: {key} differs at pairs {bad[:5]}: got {got[key][bad[:5]]} want {want[key][bad[:5]]}"
    if scope_full:
        assert np.array_equal(got["cig_off"], want["cig_off"]), f"{what}: CIGAR run counts differ"
        assert np.array_equal(got["runs"], want["runs"]), f"{what}: CIGAR runs differ"
        if want.get("locs") is not None:
            assert np.array_equal(got["locs"], want["locs"]), f"{what}: locations differ"
    else:
        assert int(got["cig_off"][-1]) == 0
