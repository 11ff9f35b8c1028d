import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import PortOracle, build

    build()
    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import RefOracle, build, have_ref

    build()
    if not have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt .so)")
    return RefOracle()


@pytest.fixture(scope="session")
def refnm():
    from oracle.oracle import RefNucMut, build, have_ref_nucmut

    build()
    if not have_ref_nucmut():
        pytest.skip("oracle/_ref/libpanman_nucmut.so not built (no /root/reference here and no prebuilt .so)")
    return RefNucMut()


@pytest.fixture(scope="session")
def refpg():
    from oracle.oracle import RefPgOrder, build, have_ref_pgorder

    build()
    if not have_ref_pgorder():
        pytest.skip("oracle/_ref/libpanman_pgorder.so not built (no /root/reference here and no prebuilt .so)")
    return RefPgOrder()
