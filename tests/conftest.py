"""pytest configuration: markers and shared fixtures.

``-m "not gpu"`` : oracle vs. golden vectors, host logic, C-ABI symbol checks (CPU only).
``-m gpu``       : parity tests proper -- CUDA path vs. oracle / goldens through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class GoldenCase:
    """One world recorded from the unmodified reference (oracle/make_golden.py)."""

    def __init__(self, npz, name):
        self.name = name
        pre = name + "/"
        self.d = {k[len(pre):]: npz[k] for k in npz.files if k.startswith(pre)}

    def __getitem__(self, k):
        return self.d[k]

    @property
    def n_steps(self):
        return int(self.d["n_steps"])

    @property
    def status(self):
        return int(self.d["status"])

    @property
    def status_step(self):
        return int(self.d["status_step"])


def load_golden(fname):
    npz = np.load(os.path.join(GOLDEN_DIR, fname))
    return [GoldenCase(npz, str(c)) for c in npz["cases"]]


@pytest.fixture(scope="session")
def physics_golden():
    return load_golden("physics_golden.npz")


@pytest.fixture(scope="session")
def failure_golden():
    return load_golden("failure_golden.npz")
