"""tools/enc_paths.py -- the same 1 029 x 1 044 480 B batch through the hot-path entry point (zpq_encode_blocks) and through the
whole-block entry point (zpq_compress_blocks): coding-kernel time of each (zpq_last_timing)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O, zpaq_b200 as Z
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 1029
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1044480
pays = [O.gen_block(1, 777 + i, n) for i in range(nblk)]
hdr = Z.model_from_method("5", pays[0])
ctx = Z.Context(0)
for rep in range(2):
    res = ctx.encode_blocks([hdr] * nblk, [b"\0" + p for p in pays])
    tm = ctx.timing(); print("encode_blocks   rep %d: code_ms %.1f init_ms %.1f  MB/s %.1f" % (rep, tm["code_ms"], tm["init_ms"], nblk * n / 1e3 / tm["code_ms"]), ctx.tiers(), flush=True)
    blocks = ctx.compress_blocks("5", pays)
    tm = ctx.timing(); print("compress_blocks rep %d: code_ms %.1f init_ms %.1f  MB/s %.1f" % (rep, tm["code_ms"], tm["init_ms"], nblk * n / 1e3 / tm["code_ms"]), ctx.tiers(), flush=True)
