import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes harness over libsvo_b200.so + synthetic data)."""
    return importlib.import_module("semi-direct-visual-odometry_b200")


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle -- the checker, never the thing under test in the gpu suite."""
    import oracle
    oracle.build()
    return oracle


_PAIRS = {}


@pytest.fixture(scope="session")
def pair_cache(pkg):
    def get(index=0, n_features=500, **kw):
        key = (index, n_features, tuple(sorted((k, (tuple(v) if hasattr(v, "__len__") else v)) for k, v in kw.items())))
        if key not in _PAIRS:
            _PAIRS[key] = pkg.synth.make_pair(index=index, n_features=n_features, **kw)
        return _PAIRS[key]
    return get
