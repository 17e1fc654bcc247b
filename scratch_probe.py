# first-contact GPU probe: times the v1 kernels on a few full-size blocks
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import oraclepy as O, zpaq_b200 as Z
ctx = Z.Context(0)
for method, nblk, n in (("4", 8, 1044480), ("5", 8, 1044480), ("5", 64, 262144), ("4", 256, 262144)):
    pays = [O.gen_block(1, 777 + i, n) for i in range(nblk)]
    hdr = Z.model_from_method(method, pays[0])
    t = time.time()
    res = ctx.encode_blocks([hdr] * nblk, [b"\0" + p for p in pays])
    dt = time.time() - t
    tm = ctx.timing()
    print("ENC", method, nblk, n, "status", set(r[0] for r in res), "wall %.2fs" % dt, tm, "MB/s %.2f" % (nblk * n / 1e6 / (tm['code_ms'] / 1e3)), flush=True)
    coded = [r[1][0] for r in res]
    if nblk <= 8 and method == "4":
        print("  parity block0:", O.encode(hdr, b"\0" + pays[0]) == coded[0], flush=True)
    t = time.time()
    dec = ctx.decode_blocks([hdr] * nblk, [c + b"\0\0\0\0" for c in coded], caps=[n + 16] * nblk)
    tm = ctx.timing()
    print("DEC", method, nblk, n, "status", set(r[0] for r in dec), "rt ok", all(dec[i][1][0] == b"\0" + pays[i] for i in range(nblk)), tm,
          "MB/s %.2f" % (nblk * n / 1e6 / (tm['code_ms'] / 1e3)), flush=True)
