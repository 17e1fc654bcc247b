"""tools/dec_check.py -- quick parity of the production decoder against the oracle's coded bytes (methods 4 / 5 / the
17-component chain, four data kinds, ragged sizes); prints which kernel tier decoded each group."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O, zpaq_b200 as Z
ctx = Z.Context(0)
CFG = "x0,0ci1,1,1,1,1,1,1,1c256c8,0,255c16,0,255,255aw1c0,0,255,1001,255c0,1010,255m16,24"
bad = 0
for method in ("5", "4", CFG):
    for kind in (0, 1, 2, 3):
        pays = [O.gen_block(kind, 100 + i, n) for i, n in enumerate((30000, 1, 0, 517, 41000, 8192, 65536))]
        groups = {}
        for p in pays:
            groups.setdefault(Z.model_from_method(method, p), []).append(p)
        for hdr, ps in groups.items():
            datas = [b"\0" + p for p in ps]
            want = [O.encode(hdr, d) for d in datas]
            dec = ctx.decode_blocks([hdr] * len(datas), [w + b"\0\0\0\0" for w in want], caps=[len(d) + 8 for d in datas])
            ok = all(st == 0 and segs[0] == datas[i] for i, (st, segs) in enumerate(dec))
            bad += not ok
            print(method[:6], "kind", kind, "n", hdr[6], "blocks", len(datas), ctx.tiers(), "ok" if ok else "MISMATCH %s" % [d[0] for d in dec], flush=True)
print("dec_check:", "ALL OK" if not bad else "%d groups FAILED" % bad)
sys.exit(1 if bad else 0)
