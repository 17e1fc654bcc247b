"""oracle/make_golden_post.py -- regenerate tests/golden/postproc.json with the UNMODIFIED reference (oracle/_ref):
small archives of the methods whose blocks need more than the PASS prefix after the model -- stored blocks (level 2),
LZ77 and BWT with their PCOMP inverse transforms, E8E9 -- for the host post-processor (zpq_postprocess_block) and the
whole-archive path (zpq_decompress_archive).  Run in the build container after `make -C oracle ref`:
    python oracle/make_golden_post.py
"""
import base64
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refpy

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "postproc.json")
entries = []
for method in ("0", "1", "2", "3", "x4,1,4,0,2,20ci1", "x4,6,4,0,2,20,16,18ci1,1m", "x4,7ci1,1m", "x4,4ci1"):
    for kind, seed, n in ((1, 71, 3000), (3, 72, 5000), (0, 73, 1200)):
        d = refpy.gen_block(kind, seed, n)
        arc = refpy.compress_block(d, method)
        assert refpy.decompress(arc, n + 16) == d
        entries.append({"method": method, "data": {"kind": kind, "seed": seed, "n": n}, "size": len(arc),
                        "sha1": hashlib.sha1(arc).hexdigest(), "archive_b64": base64.b64encode(arc).decode()})
json.dump(entries, open(OUT, "w"), indent=0)
print("wrote", OUT, len(entries), "archives,", sum(e["size"] for e in entries), "bytes")
