"""GPU: the analysis kernels in front of the coding path (zpq_prep.cu: SHA-1, level-5 gap statistics) and the
device-resident whole-block entry point zpq_compress_blocks_dev, against hashlib, the host implementation and the
unmodified reference."""
import hashlib

import pytest

import oraclepy as O
import zpaq_b200 as Z
from util import DevBytes

pytestmark = pytest.mark.gpu


def _period_block(period, n, seed):
    """rows of `period` bytes that repeat with small changes: a strong byte-gap structure at `period`"""
    import random
    r = random.Random(seed)
    row = bytes(r.randrange(256) for _ in range(period))
    out = bytearray()
    while len(out) < n:
        row = bytes(b if r.random() > 0.02 else r.randrange(256) for b in row)
        out += row
    return bytes(out[:n])


CASES = [b"", b"a", b"abc", bytes(range(256)) * 3, O.gen_block(0, 1, 55), O.gen_block(0, 2, 56), O.gen_block(1, 3, 63),
         O.gen_block(1, 4, 64), O.gen_block(2, 5, 65), O.gen_block(2, 6, 119), O.gen_block(0, 7, 120), O.gen_block(0, 8, 127),
         O.gen_block(1, 9, 128), O.gen_block(1, 10, 4096 + 3), O.gen_block(0, 11, 70001), O.gen_block(3, 12, 30011),
         _period_block(40, 20000, 1), _period_block(300, 40000, 2), _period_block(7, 9000, 3)]


def test_sha1_and_gaps_match_the_host(ctx):
    # unaligned starts: each payload sits 0..3 bytes into its buffer
    bufs = [DevBytes(b"\x55" * (i & 3) + c) for i, c in enumerate(CASES)]
    ptrs = [b.ptr + (i & 3) for i, b in enumerate(bufs)]
    sha, gaps = ctx.analyze_dev(ptrs, [len(c) for c in CASES])
    for i, c in enumerate(CASES):
        assert sha[i] == hashlib.sha1(c).digest(), ("sha1", i, len(c))
    # the host statistics show up as the period components of the level-5 header: same header either way
    for i, c in enumerate(CASES):
        want = Z.model_from_method("5", c)
        L = Z.lib()
        import ctypes as C
        L.zpq_header_from_stats.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_int32), C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        out = C.create_string_buffer(70000); n = C.c_int(0)
        g = (C.c_int32 * 2)(*gaps[i])
        assert L.zpq_header_from_stats(b"5", len(c), g, out, 70000, C.byref(n)) == 0
        assert out.raw[:n.value] == want, ("gaps", i, gaps[i])
    assert any(g[0] for g in gaps), "the periodic cases must trigger a period model"


@pytest.mark.parametrize("method", ["4", "5"])
def test_compress_blocks_dev_equals_the_host_path_and_the_reference(ctx, method):
    pays = [O.gen_block(1, 50, 9000), O.gen_block(0, 51, 5000), _period_block(40, 12000, 9), b"", O.gen_block(2, 52, 777)]
    want = ctx.compress_blocks(method, pays)
    ins = [DevBytes(p) for p in pays]
    caps = [len(p) + len(p) // 8 + 80000 for p in pays]
    outs = [DevBytes(size=c) for c in caps]
    st, ln = ctx.compress_blocks_dev(method, [b.ptr for b in ins], [len(p) for p in pays], [b.ptr for b in outs], caps)
    assert st == [0] * len(pays)
    got = [outs[i].bytes(ln[i]) for i in range(len(pays))]
    assert got == want
    import refpy
    if refpy.available():
        for p, g in zip(pays, got):
            assert g == refpy.compress_block(p, method)


def test_scan_archive_reports_segments():
    blocks = [O.gen_block(1, 60 + i, 3000 + i) for i in range(3)]
    import refpy
    if not refpy.available():
        pytest.skip("reference not built")
    arc = b"".join(refpy.compress_block(b, "4") for b in blocks)
    segs = Z.scan_archive(arc)
    assert [s.block for s in segs] == [0, 1, 2] and all(s.components == 8 for s in segs)
    for s, b in zip(segs, blocks):
        assert s.size_hint == len(b) and s.has_sha1 == 1 and bytes(s.sha1) == hashlib.sha1(b).digest()
        assert arc[s.coded_off + s.coded_len] == 253
