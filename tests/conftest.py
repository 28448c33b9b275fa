import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Builds the in-tree libraries once (the oracle, the generator, the CUDA library if nvcc is there)."""
    import __graft_entry__ as ge
    ge.build(quiet=True)


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    return o


@pytest.fixture(scope="session")
def sweeps16():
    """First 14 VLP-16-shaped sweeps of the seeded sequence (reference ring table angles)."""
    from gpscalibration_b200 import SweepGenerator
    g = SweepGenerator(sensor=0, scene=0, seed=0xC0FFEE)
    return [g.sweep(k)[0].copy() for k in range(14)]


@pytest.fixture()
def gpu():
    from gpscalibration_b200 import LoamGpu
    h = LoamGpu(device=0)
    yield h
    h.close()


def ulp_diff(a, b):
    """Distance in units in the last place between two float32 arrays."""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)
