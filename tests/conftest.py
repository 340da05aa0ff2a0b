import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as _o

    return _o.get()


@pytest.fixture(scope="session")
def lib():
    """The product library, bound to cuda:0.  GPU tests only -- it fails loudly without a device."""
    import rust_exp_b200 as pkg

    L = pkg.load()
    L.init(0)
    return L


@pytest.fixture()
def fresh(lib):
    """Library in its default state (FAST mode, default tuning)."""
    from rust_exp_b200 import binding

    lib.set_mode(binding.MODE_FAST)
    lib.tune(0, 0, 0)
    lib.bh_count_interactions(False)
    lib.phase_timing(False)
    lib.bh_partition(0)
    yield lib
    lib.set_mode(binding.MODE_FAST)
    lib.tune(0, 0, 0)
    lib.bh_partition(0)
