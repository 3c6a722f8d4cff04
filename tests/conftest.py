import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_artifacts():
    """Build whatever native piece is missing or stale (nvcc cross-compiles without a GPU), so
    the suite also runs from a fresh checkout.  Prebuilt files (the GPU box) are left alone."""
    from monte_carlo_localization_b200 import build as b
    host_so = os.path.join(ROOT, "monte_carlo_localization_b200", "host", "libpf_host.so")
    oracle_so = os.path.join(ROOT, "oracle", "liboracle.so")
    if b.is_stale() or not os.path.exists(host_so) or not os.path.exists(oracle_so):
        import __graft_entry__ as ge
        ge.build()
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _gpu_available():
    try:
        from monte_carlo_localization_b200 import capi
        return capi.load_library().mcl_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must never pass silently on a CPU box: without a device they are errors when
    # selected with -m gpu, and are deselected by -m "not gpu".
    pass
