"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks.
``-m gpu`` runs on a B200: CUDA path through the C-ABI vs the oracle.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture
def rng():
    # the reference's seed (tests/conftest.py:47-49)
    return np.random.RandomState(300696)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
