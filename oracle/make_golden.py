"""oracle/make_golden.py -- regenerate tests/golden/ with the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python oracle/make_golden.py
The GPU box never runs this; it only reads the committed fixtures.
"""
import base64
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np
import ctypes as C
import refpy
import oraclepy as O

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

CUSTOM = "x0,0ci1,1,1,1,1,1,1,1c256c8,0,255c16,0,255,255aw1c0,0,255,1001,255c0,1010,255m16,24"
CFG_ALL = open(os.path.join(HERE, "cfg_all_components.cfg")).read()
METHODS = ["4", "5", "40", "50", "54,200,1", "5,128,0", "4,100,1", "4,60,0", CUSTOM,
           "x0,0c0,0,255,300i2,12m", "x0,0c1000,3,1255a16,1,2w3,65,26,223,20,1t8s10,20,100",
           "x1,0ci1ac0,0,1300,255,2000mt12s", "x0,0w2c0,1010,255i1m", "x0,0c0,0,511c0,0,300i3"]

# ---- static tables ----
L = refpy.lib()
ns = np.zeros(1024, np.uint8); dt2k = np.zeros(256, np.int32); dt = np.zeros(1024, np.int32)
sq = np.zeros(4096, np.uint16); st = np.zeros(32768, np.int16)
assert L.ref_tables(*[a.ctypes.data_as(C.c_void_p) for a in (ns, dt2k, dt, sq, st)]) == 0
tables = {k: hashlib.sha1(a.tobytes()).hexdigest() for k, a in (("ns", ns), ("dt2k", dt2k), ("dt", dt), ("squash", sq), ("stretch", st))}
tables["stsum"] = 3887533746   # libzpaq.cpp:1759
tables["sqsum"] = 2278286169   # libzpaq.cpp:1760
tables["ns_bytes_b64"] = base64.b64encode(ns.tobytes()).decode()
json.dump(tables, open(os.path.join(OUT, "tables.json"), "w"), indent=1)

# ---- headers: method string (+ data) -> block header bytes ----
headers = []
for kind in range(4):
    for n in (0, 1000, 70000):
        d = refpy.gen_block(kind, 11, n)
        for m in METHODS:
            arc = refpy.compress_block(d, m)
            headers.append({"method": m, "kind": kind, "seed": 11, "n": n, "header": O.parse_blocks(arc)[0].header.hex()})
json.dump(headers, open(os.path.join(OUT, "headers.json"), "w"), indent=0)

# ---- whole archives: small ones inline, large ones by SHA-1 ----
arcs = []
def add(name, data_spec, arc, how):
    e = {"name": name, "data": data_spec, "how": how, "size": len(arc), "sha1": hashlib.sha1(arc).hexdigest()}
    if len(arc) <= 6000:
        e["archive_b64"] = base64.b64encode(arc).decode()
    arcs.append(e)

for kind in range(4):
    for n in (0, 1, 300, 4000):
        d = refpy.gen_block(kind, 100 + n, n)
        for m in ("4", "5", CUSTOM):
            add("m%s_k%d_n%d" % (m[:6], kind, n), {"kind": kind, "seed": 100 + n, "n": n}, refpy.compress_block(d, m),
                {"api": "compressBlock", "method": m, "filename": "", "comment": None, "dosha1": True})
d = refpy.gen_block(1, 9, 2500)
add("fn_comment", {"kind": 1, "seed": 9, "n": 2500}, refpy.compress_block(d, "4", b"some/file.txt", b"a comment", True),
    {"api": "compressBlock", "method": "4", "filename": "some/file.txt", "comment": "a comment", "dosha1": True})
add("nosha1", {"kind": 1, "seed": 9, "n": 2500}, refpy.compress_block(d, "5", b"", None, False),
    {"api": "compressBlock", "method": "5", "filename": "", "comment": None, "dosha1": False})
for lvl, cfg in ((0, CFG_ALL), (1, None), (2, None), (3, None)):
    for nseg in (1, 3):
        for kind in (0, 2):
            d = refpy.gen_block(kind, 5, 3000)
            add("cfg%d_seg%d_k%d" % (lvl, nseg, kind), {"kind": kind, "seed": 5, "n": 3000},
                refpy.compress_config(d, cfg, level=lvl, nseg=nseg),
                {"api": "Compressor", "level": lvl, "config": "oracle/cfg_all_components.cfg" if lvl == 0 else None, "nseg": nseg})
# the SURVEY.md 8(c) known answers: seed 777 'mixed', 1 044 480 bytes
d = refpy.gen_block(1, 777, 1044480)
for m in ("4", "5", CUSTOM):
    add("kat777_%s" % m[:6], {"kind": 1, "seed": 777, "n": 1044480}, refpy.compress_block(d, m),
        {"api": "compressBlock", "method": m, "filename": "", "comment": None, "dosha1": True})
json.dump(arcs, open(os.path.join(OUT, "archives.json"), "w"), indent=0)
print("wrote", len(headers), "headers,", len(arcs), "archives; KATs:", [(a["name"], a["size"], a["sha1"][:12]) for a in arcs[-3:]])
