import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of CPU time; skipped unless PNGLOSS_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("PNGLOSS_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="slow; set PNGLOSS_SLOW=1")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Build oracle/liboracle.so (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True,
                   stdout=subprocess.DEVNULL)
