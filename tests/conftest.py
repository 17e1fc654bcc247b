import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make_ctx(kernel, dec_kernel="7"):
    import zpaq_b200 as Z
    old = {k: os.environ.get(k) for k in ("ZPQ_KERNEL", "ZPQ_DEC_KERNEL")}
    os.environ["ZPQ_KERNEL"] = kernel
    os.environ["ZPQ_DEC_KERNEL"] = dec_kernel
    try:
        return Z.Context(0)   # raises if there is no GPU: GPU tests must not silently pass
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.fixture(scope="session")
def ctx():
    """Production configuration: wavefront encoder + critical-first two-warp decoder."""
    c = _make_ctx("5")
    yield c
    c.close()


@pytest.fixture(scope="session", params=["5", "5/dec6", "5/dec8", "2", "1"], ids=["production", "dec6", "dec8", "fast", "generic"])
def kctx(request):
    """Every kernel tier: 5 = production (wavefront encoder + critical-first two-warp decoder zpq_kernels7.cu), 5/dec6 = the
    round-1 two-warp decoder (generic step lists fall back to it), 5/dec8 = the opt-in lock-step decoder with lanes = blocks
    (zpq_kernels8.cu), 2 = one-warp kernel for both directions, 1 = generic kernel."""
    k, _, d = request.param.partition("/dec")
    c = _make_ctx(k, d or "7")
    yield c
    c.close()
