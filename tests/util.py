"""Shared helpers for the tests: golden fixtures, synthetic data, archive assembly."""
import base64
import hashlib
import json
import os

import oraclepy as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def data_of(spec):
    return O.gen_block(spec["kind"], spec["seed"], spec["n"])


def archive_bytes(entry):
    return base64.b64decode(entry["archive_b64"]) if "archive_b64" in entry else None


def assemble_block(hdr, coded_segments, payload_segments, filenames, comments, dosha1=True):
    """Block framing as Compressor writes it (SURVEY.md 8b block anatomy)."""
    out = bytearray(O.TAG + b"zPQ" + bytes([1 + (hdr[6] == 0), 1]) + hdr)
    for coded, pay, fn, cm in zip(coded_segments, payload_segments, filenames, comments):
        out += b"\x01" + fn + b"\0" + cm + b"\0\0" + coded + b"\0\0\0\0"
        out += (b"\xfd" + hashlib.sha1(pay).digest()) if dosha1 else b"\xfe"
    out += b"\xff"
    return bytes(out)


def split_segments(d, nseg):
    segs, pos = [], 0
    for s in range(nseg):
        ln = len(d) - pos if s == nseg - 1 else len(d) // nseg
        segs.append(d[pos:pos + ln])
        pos += ln
    return segs
