"""GPU: whole-block API (zpq_compress_blocks / zpq_decompress_archive), segments, generic HCOMP, errors."""
import hashlib

import pytest

import oraclepy as O
import zpaq_b200 as Z
from util import load, data_of, archive_bytes, split_segments

pytestmark = pytest.mark.gpu

ARCS = load("archives.json")
CB_SMALL = [a for a in ARCS if a["how"]["api"] == "compressBlock" and "archive_b64" in a]
CB_BIG = [a for a in ARCS if a["how"]["api"] == "compressBlock" and "archive_b64" not in a]
CFG = [a for a in ARCS if a["how"]["api"] == "Compressor"]


def test_compress_blocks_equals_reference_archives(ctx):
    """Batched compressBlock: every byte of every block (tag, header, segment, coded data, SHA-1) as libzpaq writes it."""
    groups = {}
    for a in CB_SMALL:
        h = a["how"]
        groups.setdefault((h["method"], h["dosha1"]), []).append(a)
    for (method, dosha1), items in groups.items():
        pays = [data_of(a["data"]) for a in items]
        fns = [a["how"]["filename"].encode() for a in items]
        cms = [a["how"]["comment"].encode() if a["how"]["comment"] is not None else None for a in items]
        got = ctx.compress_blocks(method, pays, filenames=fns, comments=cms, dosha1=dosha1)
        for a, g in zip(items, got):
            assert g == archive_bytes(a), a["name"]


def test_decompress_archive_roundtrip_and_checksums(ctx):
    pays = [data_of(a["data"]) for a in CB_SMALL if a["how"]["dosha1"]]
    arc = b"".join(archive_bytes(a) for a in CB_SMALL if a["how"]["dosha1"])
    out, nb = ctx.decompress_archive(arc, sum(map(len, pays)) + 16)
    assert nb == len(pays) and out == b"".join(pays)
    # flip one payload-coded byte: either the coder or the SHA-1 check must object
    bad = bytearray(arc)
    blk = O.parse_blocks(arc)[5]
    pos = arc.index(blk.segments[0].coded) + len(blk.segments[0].coded) // 2
    bad[pos] ^= 0x10
    with pytest.raises(Z.ZpqError):
        ctx.decompress_archive(bytes(bad), sum(map(len, pays)) + 16)


@pytest.mark.parametrize("entry", CFG, ids=[a["name"] for a in CFG])
def test_segments_and_handwritten_models(kctx, entry):
    """Multi-segment blocks (state carries over) and models whose HCOMP is not makeConfig's
    (min/mid/max.cfg, all nine component types): device bytecode interpreter path."""
    arc = archive_bytes(entry)
    blk = O.parse_blocks(arc)[0]
    d = data_of(entry["data"])
    nseg = len(blk.segments)
    segs = [(b"\0" if i == 0 else b"") + s for i, s in enumerate(split_segments(d, nseg))]
    st, coded = kctx.encode_blocks([blk.header], [b"".join(segs)], seg_lens=[[len(s) for s in segs]])[0]
    assert st == 0 and coded == [s.coded for s in blk.segments]
    packed = [c + b"\0\0\0\0" for c in coded]
    st, dec = kctx.decode_blocks([blk.header], [b"".join(packed)], seg_lens=[[len(c) for c in packed]], caps=[len(d) + 8])[0]
    assert st == 0 and dec == segs


def test_full_size_known_answers(ctx):
    """the seed-777 1 044 480-byte block at -method 4, 5 and the 17-component custom chain."""
    for a in CB_BIG:
        d = data_of(a["data"])
        got = ctx.compress_blocks(a["how"]["method"], [d])[0]
        assert len(got) == a["size"], a["name"]
        assert hashlib.sha1(got).hexdigest() == a["sha1"], a["name"]
        out, nb = ctx.decompress_archive(got, len(d))
        assert nb == 1 and hashlib.sha1(out).hexdigest() == hashlib.sha1(d).hexdigest()


def test_per_block_errors_do_not_fail_the_batch(ctx):
    d = O.gen_block(1, 4, 3000)
    hdr = Z.model_from_method("4", d)
    good = O.encode(hdr, b"\0" + d) + b"\0\0\0\0"
    trunc = good[:len(good) // 2]
    res = ctx.decode_blocks([hdr, hdr, hdr], [good, trunc, good], caps=[len(d) + 8, len(d) + 8, 10])
    assert res[0][0] == 0 and res[0][1][0] == b"\0" + d
    assert res[1][0] in (-6, -5)          # unexpected end of file / archive corrupted
    assert res[2][0] == -4                # output buffer too small
    # ZPAQL execution error: opcode 0 as the first HCOMP instruction
    bad = bytearray(hdr); bad[hdr.index(bytes([0]), 7 + 2) + 1] = 0
    st, _ = ctx.encode_blocks([bytes(bad)], [b"\0" + d])[0]
    assert st in (-3, -1)


def test_decompress_archive_every_block_type(ctx):
    """One archive holding stored (level 2) blocks, LZ77 / BWT / E8E9 blocks with their PCOMP programs and plain
    context-model blocks, all written by the unmodified reference: modelled blocks are decoded on the GPU, the
    post-processing (host ZPAQL machine) and the stored blocks are handled on the host; SHA-1s are verified."""
    post = load("postproc.json")
    arcs = [archive_bytes(e) for e in post] + [archive_bytes(a) for a in CB_SMALL[:6] if a["how"]["dosha1"]]
    pays = [data_of(e["data"]) for e in post] + [data_of(a["data"]) for a in CB_SMALL[:6] if a["how"]["dosha1"]]
    out, nb = ctx.decompress_archive(b"".join(arcs), sum(map(len, pays)) + 16)
    assert nb == len(pays) and out == b"".join(pays)
