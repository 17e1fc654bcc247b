"""tools/enc_variants.py [nblk] [bytes] -- time the wavefront encoder variants (ZPQ_WAVE_VARIANT) on one batch,
check that every variant produces the same coded bytes and that two blocks equal the CPU oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O, zpaq_b200 as Z
if os.environ.get("ZPQ_AB_LIB"):          # A/B against a library built from another commit (tools only; the package has no switch)
    Z.LIB_PATH = os.path.abspath(os.environ["ZPQ_AB_LIB"]); print("library:", Z.LIB_PATH, flush=True)

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 1029
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
variants = [int(v) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else "1,9,25,29".split(","))]
method = sys.argv[4] if len(sys.argv) > 4 else "5"
pays = [O.gen_block(1, 777 + i, n) for i in range(nblk)]
hdr = Z.model_from_method(method, pays[0])
datas = [b"\0" + p for p in pays]
ref = None
for v in variants:
    os.environ["ZPQ_WAVE_VARIANT"] = str(v)
    ctx = Z.Context(0)
    for rep in range(2):
        res = ctx.encode_blocks([hdr] * nblk, datas)
        tm = ctx.timing()
        print("variant %d rep %d: status %s code_ms %.1f  MB/s %.1f" % (v, rep, sorted(set(r[0] for r in res)), tm["code_ms"], nblk * n / 1e3 / tm["code_ms"]), flush=True)
    outs = [r[1][0] for r in res]
    if ref is None:
        ref = outs
        for i in (0, nblk - 1):
            assert outs[i] == O.encode(hdr, datas[i]), "block %d differs from the oracle" % i
        print("oracle check ok", flush=True)
    else:
        assert outs == ref, "variant %d differs" % v
        print("variant %d identical to variant %d" % (v, variants[0]), flush=True)
    ctx.close()
