import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "avisynth-sangnom2_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything the tests load is built in-tree by __graft_entry__.build(); build on demand here."""
    need = [os.path.join(PKG, n) for n in ("libsangnom_cuda.so", "libsangnom2_b200.so")]
    need.append(os.path.join(ROOT, "tests", "fakehost_src", "libfakeavs.so"))
    need.append(os.path.join(ROOT, "oracle", "liboracle.so"))
    if not all(os.path.exists(p) for p in need) and not os.environ.get("SANGNOM_SKIP_BUILD"):
        import __graft_entry__
        __graft_entry__.build()
    yield
