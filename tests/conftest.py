import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make_ctx(kernel):
    import zpaq_b200 as Z
    old = os.environ.get("ZPQ_KERNEL")
    os.environ["ZPQ_KERNEL"] = kernel
    try:
        return Z.Context(0)   # raises if there is no GPU: GPU tests must not silently pass
    finally:
        if old is None:
            del os.environ["ZPQ_KERNEL"]
        else:
            os.environ["ZPQ_KERNEL"] = old


@pytest.fixture(scope="session")
def ctx():
    """Production configuration: wavefront encoder + fast sequential decoder."""
    c = _make_ctx("5")
    yield c
    c.close()


@pytest.fixture(scope="session", params=["5", "4", "3", "2", "1"], ids=["wave", "team", "pipeline", "fast", "generic"])
def kctx(request):
    """Every kernel tier: 5 = production (wavefront encoder), 4 = team encoder, 3 = one-warp pipeline encoder, 2 = fast sequential kernel for both directions, 1 = generic kernel."""
    c = _make_ctx(request.param)
    yield c
    c.close()
