import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_tiny():
    path = os.path.join(GOLDEN_DIR, "tiny_hotpath.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_tower():
    return dict(np.load(os.path.join(GOLDEN_DIR, "tiny_tower.npz")))


@pytest.fixture(scope="session")
def tiny_problem():
    from mvsnet_b200 import synthetic
    return synthetic.make_problem("tiny")


@pytest.fixture(scope="session")
def small_problem():
    from mvsnet_b200 import synthetic
    return synthetic.make_problem("small")


def to_dev(a, dtype=None):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)
