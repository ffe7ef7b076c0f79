"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture
def no_pair_list_cache():
    """Tests of the pair-list SIZING logic (capacities, slot geometry, repeats) need every pass to build its lists."""
    from mythos_b200.energy import functional

    old = functional.PAIR_LIST_CACHE_GB
    functional.PAIR_LIST_CACHE_GB = 0.0
    functional._PAIR_LISTS.clear()
    try:
        yield
    finally:
        functional.PAIR_LIST_CACHE_GB = old
        functional._PAIR_LISTS.clear()
