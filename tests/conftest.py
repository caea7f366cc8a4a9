import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN_DIR, "reference_small.npz"))


@pytest.fixture(scope="session")
def golden_pubmed():
    return np.load(os.path.join(GOLDEN_DIR, "reference_pubmed_shape.npz"))


@pytest.fixture(scope="session")
def golden_betweenness():
    return np.load(os.path.join(GOLDEN_DIR, "reference_betweenness.npz"))


@pytest.fixture(scope="session")
def golden_eigenvector():
    return np.load(os.path.join(GOLDEN_DIR, "reference_eigenvector.npz"))


@pytest.fixture(scope="session")
def golden_deep_hub():
    return np.load(os.path.join(GOLDEN_DIR, "reference_deep_hub.npz"))


DEEP_HUB_NAMES = ("deep_chain", "hub_star", "lollipop")


def micro_names(golden):
    return sorted({k.split("/")[1] for k in golden.files if k.startswith("micro/")})
