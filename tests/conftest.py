import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ("bbcb", "heavy", "tile2", "pep28")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def atominfo_from(topo, prefix="in_"):
    info = np.empty((len(topo[prefix + "names"]), 3), dtype=object)
    info[:, 0] = [str(s) for s in topo[prefix + "names"]]
    info[:, 1] = [str(s) for s in topo[prefix + "resnames"]]
    info[:, 2] = [int(r) for r in topo[prefix + "resids"]]
    return info


@pytest.fixture(scope="session")
def param_dir():
    from molearn_b200.build import param_dir as pd
    return pd()


@pytest.fixture(scope="session", params=CASES)
def case(request):
    return request.param


@pytest.fixture(scope="session")
def golden_topo():
    cache = {}

    def get(c):
        if c not in cache:
            cache[c] = load_golden("topo_" + c)
        return cache[c]
    return get


@pytest.fixture(scope="session")
def golden_energy():
    cache = {}

    def get(c):
        if c not in cache:
            cache[c] = load_golden("energy_" + c)
        return cache[c]
    return get


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
