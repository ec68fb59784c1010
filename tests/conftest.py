import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "accelerated-ray-tracer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_gpu")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def texture_dir():
    for d in (os.path.join(ROOT, "oracle", "_ref", "textures"), os.path.join(ROOT, "tests", "golden", "textures")):
        if os.path.exists(os.path.join(d, "earthmap.ppm")):
            return d
    return None


@pytest.fixture(scope="session")
def built():
    """Everything native is built in-tree by __graft_entry__.build() (a no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope="session")
def pyrt(built):
    import pyrt as m
    return m


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        p = os.path.join(GOLDEN, name + ".npz")
        if not os.path.exists(p):
            pytest.skip("golden %s missing" % name)
        return np.load(p)
    return load
