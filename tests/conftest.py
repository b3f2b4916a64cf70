"""pytest configuration: `gpu` marker + shared fixtures.

CPU suite (`-m "not gpu"`): oracle vs mathematics / RFC vectors / committed golden fixtures,
host-side logic, and that the C-ABI library loads and exports every symbol of include/*.h.
GPU suite (`-m gpu`): the CUDA path, called through the C ABI, against the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "lib", "libmpshuffle.so")):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)  # raises if there is no CUDA device: no CPU fallback
    yield c
    c.close()
