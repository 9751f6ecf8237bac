import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    o.load()
    return o


@pytest.fixture(scope="session")
def mitr():
    import mitransient_b200
    return mitransient_b200


# Every GPU test runs three times: with the library's own choice of histogram path (TR_HIST_AUTO) and with each of the
# two paths preferred (shared-memory slots / global reductions), so that toy-size inputs cover both kernels' code.
def pytest_generate_tests(metafunc):
    if metafunc.definition.get_closest_marker("gpu") is not None:
        metafunc.parametrize("hist_path", ["auto", "shared", "global"], indirect=True)


@pytest.fixture(autouse=True)
def hist_path(request, monkeypatch):
    which = getattr(request, "param", "auto")
    monkeypatch.setenv("MITRANSIENT_B200_HISTOGRAM_PATH", which)   # read by every integrator built during the test
    return which
