import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a kernel that never returns must fail the test, not hang the box (pytest-timeout)
    for item in items:
        if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(240))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle as O  # oracle/oracle.py (test infrastructure)
    O.build()
    return O


GOLDEN_SCENES = ["wavefront_comparison", "cornell_smoke", "feature_zoo", "vol2_final_scene"]
