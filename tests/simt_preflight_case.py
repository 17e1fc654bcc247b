"""Run by tests/test_simt_preflight.py through tools/simt/run.py: a small encode / decode of every production kernel
path on the SIMT emulator build, bit for bit against the oracle.  (Not collected by pytest: no test_ prefix.)"""
import oraclepy as O
import zpaq_b200 as Z

CFG = "x0,0ci1,1,1,1,1,1,1,1c256c8,0,255c16,0,255,255aw1c0,0,255,1001,255c0,1010,255m16,24"
ctx = Z.Context(0)
checked = 0
for method in ("4", "5", CFG):
    pays = [O.gen_block(kind, 40 + kind, n) for kind, n in ((1, 2600), (3, 1500), (0, 0), (2, 700))]
    groups = {}
    for p in pays:
        groups.setdefault(Z.model_from_method(method, p), []).append(p)
    for hdr, ps in groups.items():
        datas = [b"\0" + p for p in ps]
        want = [O.encode(hdr, d) for d in datas]
        enc = ctx.encode_blocks([hdr] * len(datas), datas)
        assert all(st == 0 and segs[0] == want[i] for i, (st, segs) in enumerate(enc)), ("encode", method)
        dec = ctx.decode_blocks([hdr] * len(datas), [w + b"\0\0\0\0" for w in want], caps=[len(d) + 8 for d in datas])
        assert all(st == 0 and segs[0] == datas[i] for i, (st, segs) in enumerate(dec)), ("decode", method)
        checked += len(datas)
blocks = ctx.compress_blocks("5", [O.gen_block(1, 7, 1800), O.gen_block(3, 8, 900)])
out, nb = ctx.decompress_archive(b"".join(blocks), 2700)
assert nb == 2 and out == O.gen_block(1, 7, 1800) + O.gen_block(3, 8, 900)
assert ctx.launch_count() > 0
ctx.close()

# pre-processing kernels (zpq_bwt.cu, zpq_lz77.cu): BWT, E8E9, LZ77, store against blocks written by the unmodified reference
import hashlib
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import transform_payload
gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transforms.json")))["cases"]
ctx = Z.Context(0)
by_method = {}
PRE = {"x0,3ci1": 12000, "x0,7ci1": 12000, "x5,3ci1": 12000, "x5,7ci1": 12000, "x0,4ci1,1,1,1,2am": 12000,        # BWT, E8E9
       "0": 9000, "x0,0": 70000, "1": 30000, "x0,2,4,0,3,20": 20000, "x0,6,8,0,3,18c0,0,511": 20000, "2": 30000}     # store, LZ77 (hash, suffix array)
for c in gold:
    if c["n"] <= PRE.get(c["method"], -1):
        by_method.setdefault(c["method"], []).append(c)
nt = 0
for method, cs in by_method.items():
    blocks = ctx.compress_blocks(method, [transform_payload(c) for c in cs])
    for c, b in zip(cs, blocks):
        assert len(b) == c["block_len"] and hashlib.sha1(b).hexdigest() == c["block_sha1"], ("transform", method, c["n"])
        nt += 1
    out, nb = ctx.decompress_archive(b"".join(blocks), sum(c["n"] for c in cs))
    assert nb == len(cs) and out == b"".join(transform_payload(c) for c in cs), ("transform round trip", method)
ctx.close()
print("simt preflight ok: %d blocks, %d transformed blocks" % (checked, nt))
