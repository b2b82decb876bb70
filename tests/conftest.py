"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here on CPU (oracle vs golden vectors and vs the compiled reference, host logic, symbol
export checks, gloo multi-process sharding); `-m gpu` runs on a B200 and calls the CUDA path through the C-ABI.
Nothing in the gpu tests reads /root/reference.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def po():
    """The CPU oracle front end (test infrastructure)."""
    mod = ge.load_oracle()
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        ge.build()
    return mod


@pytest.fixture(scope="session")
def oracle(po):
    return po.Oracle()


@pytest.fixture(scope="session")
def reflib(po):
    if not po.RefLib.available():
        pytest.skip("oracle/_ref/librscape_ref.so not built (needs /root/reference at build time)")
    return po.RefLib()


@pytest.fixture(scope="session")
def pkg():
    mod = ge.load_package()
    if not os.path.exists(mod.LIB_PATH):
        ge.build()
    return mod


@pytest.fixture()
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()
