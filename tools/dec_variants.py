"""tools/dec_variants.py [nblk] [bytes] [variants] [method] -- time the two-warp decoder variants (ZPQ_DEC_VARIANT) on one
batch of blocks coded by the GPU encoder; every variant must reproduce the input."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O, zpaq_b200 as Z

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 1029
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
variants = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "3,7,15,31").split(",")]
method = sys.argv[4] if len(sys.argv) > 4 else "5"
pays = [O.gen_block(1, 777 + i, n) for i in range(nblk)]
hdr = Z.model_from_method(method, pays[0])
datas = [b"\0" + p for p in pays]
ctx = Z.Context(0)
enc = ctx.encode_blocks([hdr] * nblk, datas)
assert all(r[0] == 0 for r in enc)
assert enc[0][1][0] == O.encode(hdr, datas[0]), "encoder differs from the oracle"
coded = [r[1][0] + b"\0\0\0\0" for r in enc]
ctx.close()
for v in variants:
    os.environ["ZPQ_DEC_VARIANT"] = str(v)
    ctx = Z.Context(0)
    for rep in range(2):
        dec = ctx.decode_blocks([hdr] * nblk, coded, caps=[len(d) + 8 for d in datas])
        tm = ctx.timing()
        ok = all(st == 0 and segs[0] == datas[i] for i, (st, segs) in enumerate(dec))
        print("dec variant %d rep %d: %s code_ms %.1f  MB/s %.1f" % (v, rep, "ok" if ok else "MISMATCH", tm["code_ms"], nblk * n / 1e3 / tm["code_ms"]), flush=True)
    ctx.close()
