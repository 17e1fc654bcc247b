#!/usr/bin/env python
"""Golden vectors for the pre-processing paths (BWT, E8E9): blocks written by the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/Makefile) -> tests/golden/transforms.json.  Run where /root/reference exists:
    python oracle/make_golden_transforms.py
Every entry: method string, payload recipe (tests/util.py: transform_payload), the block header bytes, the bytes that pass
through the model in front of the transformed payload (the PCOMP program; recovered by decoding the reference's block with
the oracle), length and SHA-1 of the whole block."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import oraclepy as O
import refpy as R
from util import transform_payload

CASES = [
    # method, kind, seed, n, variant
    ("x0,3ci1", 0, 11, 20000, "plain"), ("x0,3ci1", 1, 12, 33333, "plain"), ("x0,3ci1", 2, 13, 4099, "plain"),
    ("x0,3ci1", 0, 14, 0, "plain"), ("x0,3ci1", 0, 15, 1, "plain"), ("x0,3ci1", 3, 16, 9000, "plain"),
    ("x0,7ci1", 2, 17, 30000, "exe"), ("x0,7ci1", 2, 18, 7, "exe"),
    ("x5,3ci1", 0, 19, 12000, "plain"), ("x5,7ci1", 2, 20, 12000, "exe"),
    ("x0,4ci1,1,1,1,2am", 2, 21, 30000, "exe"), ("x0,4ci1,1,1,1,2am", 2, 22, 4, "exe"), ("x0,4ci1,1,1,1,2am", 2, 22, 5, "exe"),
    ("3,200,1", 0, 23, 25000, "plain"), ("3,165,0", 1, 24, 25000, "plain"), ("3,200,3", 2, 25, 25000, "exe"),
    ("4,230,0", 0, 26, 25000, "plain"), ("4,230,2", 2, 27, 25000, "exe"), ("4,100,2", 2, 28, 25000, "exe"), ("4,100,3", 2, 29, 25000, "exe"),
    ("5,100,2", 2, 30, 25000, "exe"), ("5,0,3", 2, 31, 25000, "exe"),
    ("x0,3c0,0,255i2m", 0, 32, 15000, "plain"),
    # store and LZ77 selections (LZBuffer levels 1 / 2: hash table and suffix array variants, with E8E9, with a model on top)
    ("0", 0, 40, 9000, "plain"), ("x0,0", 1, 41, 70000, "plain"), ("1,5,0", 2, 42, 9000, "plain"),
    ("1", 0, 43, 30000, "plain"), ("1", 0, 43, 0, "plain"), ("1", 0, 43, 3, "plain"), ("1,15,0", 1, 44, 30000, "plain"), ("1,25,2", 2, 45, 30000, "exe"),
    ("1,40,0", 3, 46, 30000, "plain"), ("1,250,3", 2, 47, 30000, "exe"),
    ("2", 0, 48, 30000, "plain"), ("2,10,0", 1, 49, 30000, "plain"), ("2,100,2", 2, 50, 20000, "exe"),
    ("3", 1, 51, 30000, "plain"), ("3,10,3", 2, 52, 20000, "exe"), ("3,100,2", 2, 53, 20000, "exe"),
    ("4,5,0", 0, 54, 20000, "plain"), ("4,10,0", 0, 55, 20000, "plain"), ("4,10,2", 2, 56, 20000, "exe"),
    ("x0,2,4,0,3,20", 0, 57, 20000, "plain"), ("x0,6,8,0,3,18c0,0,511", 2, 58, 20000, "exe"), ("x5,1,4,0,3,24", 0, 59, 20000, "plain"),
    ("x0,2,12,0,7,21,1c0,0,511i2", 3, 60, 40000, "plain"),
]
REJECTED = ["x0,8ci1", "x0,1,3,0,3,24", "x0,2,0,0,3,20", "y0"]        # no such pre-processing; minimum match too small; not a method


def main():
    out = {"cases": [], "rejected": REJECTED}
    for method, kind, seed, n, variant in CASES:
        spec = {"method": method, "kind": kind, "seed": seed, "n": n, "variant": variant}
        p = transform_payload(spec)
        blk = R.compress_block(p, method)
        b = O.parse_blocks(blk)[0]
        if b.header[6] > 0:
            model_in = O.decode(b.header, b.segments[0].coded + b"\0\0\0\0", n + 70000)
        else:                                   # stored: 64 KiB chunks behind 4-byte lengths
            c, model_in, at = b.segments[0].coded, b"", 0
            while at + 4 <= len(c):
                ln = int.from_bytes(c[at:at + 4], "big")
                model_in += c[at + 4:at + 4 + ln]
                at += 4 + ln
        plen = 1 if model_in[0] == 0 else 3 + model_in[1] + 256 * model_in[2]
        spec.update(header=b.header.hex(), prefix=model_in[:plen].hex(), model_bytes=len(model_in), block_len=len(blk), block_sha1=hashlib.sha1(blk).hexdigest())
        out["cases"].append(spec)
        assert R.decompress(blk, n + 16) == p
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "transforms.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
