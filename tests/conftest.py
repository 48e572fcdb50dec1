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


@pytest.fixture
def tuning():
    """Set development switches of the library for one test (mvsb200_set_tuning) and restore the defaults after it."""
    from mvsnet_b200 import _lib
    touched = []

    def set_(name, value):
        _lib.set_tuning(name, value)
        touched.append(name)

    yield set_
    for name in touched:
        _lib.set_tuning(name, None)


def oracle_hot_path(problem, threads=None):
    """The CPU oracle over a whole problem with the warp + variance spread over a thread pool (numpy releases the GIL):
    the same functions as oracle.inference_from_features, fast enough for BASELINE config 1 / 2 once per session."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    import oracle as O
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    feats, cams = problem["feats"], problem["cams"]
    ds, di, D = problem["depth_start"], problem["depth_interval"], problem["depth_num"]
    n = feats.shape[0]
    H = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], D, ds, di)[0] for v in range(1, n)])
    with ThreadPoolExecutor(max_workers=threads) as ex:
        cost = np.stack(list(ex.map(lambda d: O.cost_volume(feats, H[:, d:d + 1])[0], range(D))))
    filtered = O.regnet_us0(cost, problem["weights"])
    depth, prob, _P = O.depth_regress(filtered, ds, di)
    return dict(homographies=H, cost=cost, filtered=filtered, depth=depth, prob=prob)


@pytest.fixture(scope="session")
def oracle_cfg1():
    from mvsnet_b200 import synthetic
    p = synthetic.make_problem("cfg1")
    return p, oracle_hot_path(p)


@pytest.fixture(scope="session")
def oracle_cfg2():
    from mvsnet_b200 import synthetic
    p = synthetic.make_problem("cfg2")
    r = oracle_hot_path(p)
    r.pop("cost")             # 1.5 GB: the tests that need planes of it recompute them
    return p, r
