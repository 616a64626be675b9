import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_files():
    import glob
    files = [f for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
             if not os.path.basename(f).startswith(("road_", "ingest_"))]   # those: test_road_oracle.py, test_ingest.py
    assert files, "tests/golden/*.npz missing"
    return files
