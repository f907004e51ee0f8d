"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 box.

Only tests/ (plus __graft_entry__.smoke and bench.py's CPU legs) may touch oracle/ — the oracle is the
checker, never the thing shipped.
"""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(directory, target):
    # Always: make is incremental, and a prebuilt .so that is older than an edited source must not pass the tests
    # (the .so files are git-ignored and travel to the GPU box with the snapshot).
    subprocess.run(["make", "-C", directory, "-j8"], check=True, stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(directory, target))


@pytest.fixture(scope="session")
def pkg():
    _make(os.path.join(ROOT, "zig-raytracing-weekend_b200"), "_lib/librtw_host.so")
    return importlib.import_module("zig-raytracing-weekend_b200")


@pytest.fixture(scope="session")
def orc():
    _make(os.path.join(ROOT, "oracle"), "liboracle.so")
    import oracle_ffi
    return oracle_ffi


@pytest.fixture(scope="session")
def earthmap():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "earthmap_rgb.npz")
    rgb = np.load(path)["rgb"]
    rgba = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2)
    return np.ascontiguousarray(rgba)
