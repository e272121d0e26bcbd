import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """libmvo.so + oracle built in-tree (the GPU box uses the prebuilt files that travel with the snapshot)."""
    import __graft_entry__ as g
    lib = ROOT / "monocular-visual-odometry_b200" / "libmvo.so"
    orc = ROOT / "oracle" / "build" / "liboracle.so"
    if not lib.exists() or not orc.exists():
        g.build()
    return lib


@pytest.fixture(scope="session")
def ctx(built):
    import mvo_b200
    c = mvo_b200.Context(0)     # raises if there is no B200: GPU tests must not pass on a fallback
    yield c
    c.close()


def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False
