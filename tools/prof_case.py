"""tools/prof_case.py METHOD NBLK NBYTES [enc|dec|both] [kind] -- one small batch for ncu / timing."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import oraclepy as O, zpaq_b200 as Z
method, nblk, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "both"
kind = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ctx = Z.Context(0)
pays = [O.gen_block(kind, 777 + i, n) for i in range(nblk)]
hdrs = [Z.model_from_method(method, p) for p in pays]
res = ctx.encode_blocks(hdrs, [b"\0" + p for p in pays])
tm = ctx.timing()
print("ENC", method, nblk, n, set(r[0] for r in res), tm, "MB/s %.3f" % (nblk * n / 1e3 / tm['code_ms']), "us/bit/block %.3f" % (tm['code_ms'] * 1e3 / (n * 8 + 8)), flush=True)
if mode in ("dec", "both"):
    coded = [r[1][0] for r in res]
    dec = ctx.decode_blocks(hdrs, [c + b"\0\0\0\0" for c in coded], caps=[n + 16] * nblk)
    tm = ctx.timing()
    print("DEC", method, nblk, n, set(r[0] for r in dec), all(dec[i][1][0] == b"\0" + pays[i] for i in range(nblk)), tm,
          "MB/s %.3f" % (nblk * n / 1e3 / tm['code_ms']), "us/bit/block %.3f" % (tm['code_ms'] * 1e3 / (n * 8 + 8)), flush=True)
