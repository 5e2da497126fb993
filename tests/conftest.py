import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200")
for p in (os.path.join(PKG, "python"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    """Build the product library and the oracle once per session (nvcc cross-compiles without a GPU)."""
    import subprocess
    if not os.path.exists(os.path.join(PKG, "libicp_b200.so")) or not os.path.exists(os.path.join(PKG, "apps", "selftest_my_lib")):
        subprocess.run(["make", "-C", PKG, "all"], check=True, stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True, stdout=subprocess.DEVNULL)


_ensure_built()


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle


@pytest.fixture(scope="session")
def ib():
    import icp_b200
    return icp_b200


@pytest.fixture(scope="session")
def ctx(ib):
    c = ib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
