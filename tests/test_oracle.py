"""CPU: the oracle (oracle/zpq_oracle.c) against the reference's own known answers and the golden
vectors made with the unmodified reference (oracle/make_golden.py)."""
import hashlib

import numpy as np
import pytest

import oraclepy as O
import zpaq_b200 as Z
from util import load, data_of, archive_bytes, assemble_block, split_segments


def test_table_known_answers():
    gold = load("tables.json")
    st, sq = O.table_sums()
    assert st == gold["stsum"] == 3887533746      # libzpaq.cpp:1759
    assert sq == gold["sqsum"] == 2278286169      # libzpaq.cpp:1760
    t = O.tables()
    for k in ("ns", "dt2k", "dt", "squash", "stretch"):
        assert hashlib.sha1(t[k].tobytes()).hexdigest() == gold[k], k


ARCS = load("archives.json")
SMALL = [a for a in ARCS if "archive_b64" in a]
BIG = [a for a in ARCS if "archive_b64" not in a]


@pytest.mark.parametrize("entry", SMALL, ids=[a["name"] for a in SMALL])
def test_oracle_matches_reference_archives(entry):
    arc = archive_bytes(entry)
    assert hashlib.sha1(arc).hexdigest() == entry["sha1"]
    blk = O.parse_blocks(arc)[0]
    d = data_of(entry["data"])
    nseg = len(blk.segments)
    pay = split_segments(d, nseg)
    segs = [(b"\0" if i == 0 else b"") + s for i, s in enumerate(pay)]
    coded = O.encode_segs(blk.header, segs)
    assert coded == [s.coded for s in blk.segments]
    dec = O.decode_segs(blk.header, [c + b"\0\0\0\0" for c in coded], len(d) + 1)
    assert dec == segs
    rebuilt = assemble_block(blk.header, coded, pay, [s.filename for s in blk.segments], [s.comment for s in blk.segments],
                             dosha1=blk.segments[0].sha1 is not None)
    assert rebuilt == arc


@pytest.mark.parametrize("entry", BIG, ids=[a["name"] for a in BIG])
def test_oracle_full_size_known_answers(entry):
    """SURVEY.md 8(c) KATs: seed-777 'mixed' block of 1 044 480 B; header from the product's generator."""
    d = data_of(entry["data"])
    hdr = Z.model_from_method(entry["how"]["method"], d)
    coded = O.encode(hdr, b"\0" + d)
    arc = assemble_block(hdr, [coded], [d], [b""], [str(len(d)).encode()])
    assert len(arc) == entry["size"]
    assert hashlib.sha1(arc).hexdigest() == entry["sha1"]


def test_oracle_rejects_corruption():
    d = O.gen_block(1, 3, 2000)
    hdr = Z.model_from_method("4", d)
    coded = bytearray(O.encode(hdr, b"\0" + d))
    with pytest.raises(O.OracleError):
        O.decode(hdr, bytes(coded[:len(coded) // 2]), len(d) + 1)       # truncated: unexpected end of file
    coded[10] ^= 0x55
    try:
        out = O.decode(hdr, bytes(coded) + b"\0\0\0\0", len(d) + 1)
        assert out != b"\0" + d
    except O.OracleError:
        pass
