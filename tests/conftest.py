import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "msda-triton_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _library_knobs_follow_the_environment():
    """libmsda_b200 reads its MSDA_B200_* knobs once; tests that flip one call util.knobs() / _lib.reload_tuning().
    After every test (monkeypatch has restored the environment by then) the library re-reads them."""
    yield
    try:
        from msda_triton import _lib
    except Exception:
        return
    if _lib._lib_handle is not None:
        _lib.reload_tuning()
