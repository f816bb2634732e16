import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def nid():
    """The product package (ctypes mirror of include/nid_b200.h)."""
    return importlib.import_module("nid-pose-estimation_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("nid-pose-estimation_b200.synth")


@pytest.fixture(scope="session")
def orc():
    from oracle import binding
    binding.lib()
    return binding


_PAIRS = {}


@pytest.fixture(scope="session")
def make_pair(synth):
    def _mk(seed=1000, rows=120, cols=160, **kw):
        key = (seed, rows, cols, tuple(sorted(kw.items())))
        if key not in _PAIRS:
            _PAIRS[key] = synth.make_pair(seed, rows, cols, **kw)
        return _PAIRS[key]
    return _mk


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor)
    return np.abs(a - b) / np.where(den > 0, den, 1.0)
