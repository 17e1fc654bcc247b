"""GPU parity for the pre-processing kernels (zpq_bwt.cu, zpq_lz77.cu): zpq_compress_blocks with a store / LZ77 / BWT / E8E9 method
must write the bytes libzpaq::compressBlock writes (LZBuffer + e8e9 + makeConfig's PCOMP, libzpaq.cpp:6450-6883, 6920-7323)."""
import hashlib

import pytest

import oraclepy as O
import zpaq_b200 as Z
from util import exe_like, load, transform_payload

pytestmark = pytest.mark.gpu
GOLD = load("transforms.json")["cases"]


def _by_method():
    d = {}
    for c in GOLD:
        d.setdefault(c["method"], []).append(c)
    return sorted(d.items())


@pytest.mark.parametrize("method,cases", _by_method(), ids=[m for m, _ in _by_method()])
def test_blocks_equal_the_reference_goldens(ctx, method, cases):
    pays = [transform_payload(c) for c in cases]
    l0 = ctx.launch_count()
    blocks = ctx.compress_blocks(method, pays)
    assert ctx.launch_count() > l0
    for c, b in zip(cases, blocks):
        assert len(b) == c["block_len"] and hashlib.sha1(b).hexdigest() == c["block_sha1"], (method, c["n"])
    out, nb = ctx.decompress_archive(b"".join(blocks), sum(len(p) for p in pays))
    assert nb == len(pays) and out == b"".join(pays)


def test_full_size_batch_against_the_live_reference(ctx):
    """48 blocks x 1 044 480 B of three kinds through BWT in one call (one index space for all suffixes), an E8E9 + BWT block and
    a block of another size class (1.5 MiB: ph = pm = 21, suffix ranks of 21 bits); some of them compared with the unmodified
    reference, seven decoded again in ONE call (a launch costs the latency of its longest block, whatever it holds).  The
    program variant of blocks above 16 MiB is covered by the `x5,...` goldens on small payloads: one block of that size alone
    is minutes of coder latency."""
    import refpy as R
    n = (1 << 20) - 4096
    pays = [O.gen_block(k % 3, 7000 + k, n) for k in range(48)]
    blocks = ctx.compress_blocks("x0,3ci1", pays)
    e = exe_like(99, n)
    be = ctx.compress_blocks("x0,7ci1", [e])[0]
    big = O.gen_block(0, 4242, 3 << 19)
    bb = ctx.compress_blocks("31,200,1", [big])[0]
    hdr = O.parse_blocks(bb)[0].header
    assert hdr[4] == 21 and hdr[5] == 21
    if R.available():
        for k in (0, 1):
            assert blocks[k] == R.compress_block(pays[k], "x0,3ci1"), k
        assert be == R.compress_block(e, "x0,7ci1")
        assert bb == R.compress_block(big, "31,200,1")
    want = b"".join(pays[:5]) + e + big
    out, nb = ctx.decompress_archive(b"".join(blocks[:5]) + be + bb, len(want))
    assert nb == 7 and out == want


def test_lz77_batch_against_the_live_reference(ctx):
    """32 blocks x 256 KiB through the LZ77 parsers (one GPU thread per block): hash-table search with variable-length codes
    (`1`), suffix-array search (`2`: the batch is suffix-sorted by the BWT's sorter first), byte codes under a model (`3,100,0`),
    and stored blocks (`1,5,0`); the first and the last block of each against the unmodified reference, all decoded again."""
    import refpy as R
    n = 1 << 18
    pays = [O.gen_block(k % 4, 9000 + k, n - 37 * k) for k in range(32)]
    for method in ("1", "2", "3,100,0", "1,5,0"):
        blocks = ctx.compress_blocks(method, pays)
        if R.available():
            for k in (0, 31):
                assert blocks[k] == R.compress_block(pays[k], method), (method, k)
        out, nb = ctx.decompress_archive(b"".join(blocks), sum(len(p) for p in pays))
        assert nb == 32 and out == b"".join(pays), method


def test_degenerate_payloads(ctx):
    """long runs and long repeats: many doubling rounds, groups that stay large"""
    import refpy as R
    pays = [b"\0" * 300000, b"abcd" * 60000, bytes(range(256)) * 500, b"a" * 70000 + b"b" + b"a" * 70000]
    blocks = ctx.compress_blocks("x0,3ci1", pays)
    if R.available():
        for p, b in zip(pays, blocks):
            assert b == R.compress_block(p, "x0,3ci1")
    out, nb = ctx.decompress_archive(b"".join(blocks), sum(len(p) for p in pays))
    assert out == b"".join(pays)
