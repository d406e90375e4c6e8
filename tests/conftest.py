import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_library():
    """(Re)build libalrender.so when nvcc is around and the sources are newer than the binary."""
    try:
        from audiblelight_b200 import build
        if build.is_stale():
            build.build_library()
    except Exception as exc:  # no nvcc: tests that need the library skip or fail loudly on their own
        print(f"[conftest] libalrender.so not (re)built: {exc}")


def pytest_configure(config):
    _ensure_library()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
