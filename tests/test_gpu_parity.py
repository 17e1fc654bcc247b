"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import pytest

import oraclepy as O
import zpaq_b200 as Z

pytestmark = pytest.mark.gpu

CFG_ALL_METHOD = "x0,0ci1,1,1,1,1,1,1,1c256c8,0,255c16,0,255,255aw1c0,0,255,1001,255c0,1010,255m16,24"


def _roundtrip(ctx, hdr, payloads):
    datas = [b"\0" + p for p in payloads]
    want = [O.encode(hdr, d) for d in datas]
    got = ctx.encode_blocks([hdr] * len(datas), datas)
    for i, (st, segs) in enumerate(got):
        assert st == 0, (i, st)
        assert segs[0] == want[i], "coded bytes differ from the oracle in block %d" % i
    dec = ctx.decode_blocks([hdr] * len(datas), [w + b"\0\0\0\0" for w in want], caps=[len(d) + 16 for d in datas])
    for i, (st, segs) in enumerate(dec):
        assert st == 0, (i, st)
        assert segs[0] == datas[i]


@pytest.mark.parametrize("method", ["4", "5", CFG_ALL_METHOD])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_method_parity(kctx, method, kind):
    ctx = kctx
    payloads = [O.gen_block(kind, 100 + i, n) for i, n in enumerate((0, 1, 7, 300, 5000, 40000))]
    hdrs = {}
    for p in payloads:
        hdrs.setdefault(Z.model_from_method(method, p), []).append(p)
    for hdr, ps in hdrs.items():
        _roundtrip(ctx, hdr, ps)


@pytest.mark.parametrize("env,values", [("ZPQ_DEC_VARIANT", ["0", "1", "2", "3", "7", "15", "31"]), ("ZPQ_WAVE_VARIANT", ["0", "1", "2", "3", "4", "5", "8", "9", "24", "25"])])
def test_kernel_variants_are_bit_exact(env, values):
    """Every switchable variant of the production kernels (two-warp decoder: candidate entries / rows, slot-scheduled
    hashes, MATCH in warp A, register-resident predictions in warp B; wavefront encoder: byte-ahead MIX rows, split launches, mixers as component threads, byte-batched MIX warps, nibble-unrolled ICM / ISSE loop) codes the same bytes as the oracle."""
    import os
    payloads = [O.gen_block(kind, 500 + kind, 6000 + 111 * kind) for kind in (0, 1, 2, 3)]
    for v in values:
        old = os.environ.get(env)
        os.environ[env] = v
        try:
            ctx = Z.Context(0)
        finally:
            if old is None:
                del os.environ[env]
            else:
                os.environ[env] = old
        try:
            for method in ("4", "5", CFG_ALL_METHOD):
                hdrs = {}
                for p in payloads:
                    hdrs.setdefault(Z.model_from_method(method, p), []).append(p)
                for hdr, ps in hdrs.items():
                    _roundtrip(ctx, hdr, ps)
        finally:
            ctx.close()
