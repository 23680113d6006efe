import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc  # oracle/oracle.py -- the CPU checker (test infrastructure)
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def pkg():
    import mgicp_b200
    return mgicp_b200


@pytest.fixture(scope="session")
def engine(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg.Engine(0)


@pytest.fixture(scope="session")
def pair30k(pkg):
    """config 1 of BASELINE.json: ~30k-point synthetic HDL-32 pair with a known perturbation"""
    return pkg.synthetic.make_pair(1000, seed=0)


@pytest.fixture(scope="session")
def pair_small(pkg):
    return pkg.synthetic.make_pair(250, seed=3)


def rows_as_keys(a):
    """view each row of a float64 [n,3] array as raw bytes (bit-exact identity of points)"""
    a = np.ascontiguousarray(a, np.float64)
    return a.view([("", a.dtype)] * a.shape[1]).reshape(-1)


def sort_rows(a):
    a = np.asarray(a, np.float64)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def match_rows(a, b):
    """index array idx with a[i] == b[idx[i]] bit-exactly (both must hold the same set of distinct rows)"""
    ka, kb = rows_as_keys(a), rows_as_keys(b)
    ob = np.argsort(kb)
    pos = np.searchsorted(kb[ob], ka)
    pos = np.clip(pos, 0, len(ob) - 1)
    idx = ob[pos]
    assert np.array_equal(kb[idx], ka), "point sets differ"
    return idx
