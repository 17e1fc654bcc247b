"""CPU tier for SURVEY.md 8(f)-4: what a method that pre-processes the payload (E8E9, BWT) puts around the model --
header bytes (ph / pm announce the post-processor) and the PCOMP program that passes through the model first -- against
blocks written by the unmodified reference (tests/golden/transforms.json, oracle/make_golden_transforms.py).
The transforms themselves are CUDA kernels: tests/test_gpu_transforms.py, and the emulator pre-flight."""
import pytest

import zpaq_b200 as Z
from util import load, transform_payload

GOLD = load("transforms.json")


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "%s-%d" % (c["method"], c["n"]))
def test_plan_bytes_equal_the_reference(case):
    hdr, prep, prefix = Z.block_plan(case["method"], case["n"])
    assert hdr.hex() == case["header"]
    assert prefix.hex() == case["prefix"]
    want = case["method"].split(",")[1][:1] if case["method"].startswith("x") and "," in case["method"] else None
    if want is not None:
        assert prep == int(want)
    assert (prefix[0] == 0) == (prep == 0)


def test_header_only_entry_point_refuses_transforming_methods():
    """zpq_model_from_method returns a header for callers that feed the model themselves: a method that transforms the payload
    cannot be honoured there."""
    with pytest.raises(Z.ZpqError):
        Z.model_from_method("x0,3ci1", b"abc" * 100)
    with pytest.raises(Z.ZpqError):
        Z.model_from_method("4,230,0", b"abc" * 100)


@pytest.mark.parametrize("method", GOLD["rejected"])
def test_methods_libzpaq_refuses_are_refused(method):
    with pytest.raises(Z.ZpqError) as e:
        Z.block_plan(method, 20000)
    assert e.value.rc == -13


def test_live_reference_agrees_with_the_goldens():
    import hashlib
    import refpy as R
    if not R.available():
        pytest.skip("reference not built here")
    for case in GOLD["cases"][:6]:
        blk = R.compress_block(transform_payload(case), case["method"])
        assert hashlib.sha1(blk).hexdigest() == case["block_sha1"]
