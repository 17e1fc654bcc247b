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


@pytest.mark.parametrize("env,values", [("ZPQ_DEC_VARIANT", ["0", "3", "15", "31"]), ("ZPQ_WAVE_VARIANT", ["0", "1", "2", "4", "9", "24", "25", "41", "57"])])
def test_kernel_variants_are_bit_exact(env, values):
    """Every switchable variant of the production kernels (two-warp decoder: candidate entries / rows, slot-scheduled
    hashes, MATCH in warp A, register-resident predictions in warp B; wavefront encoder: byte-ahead MIX rows, split launches, mixers as component threads, byte-batched MIX warps, nibble-unrolled ICM / ISSE loop) codes the same bytes as the oracle."""
    import os
    payloads = [O.gen_block(kind, 500 + kind, 6000 + 111 * kind) for kind in (0, 1, 2, 3)]
    for v in values:
        old = {k: os.environ.get(k) for k in (env, "ZPQ_DEC_KERNEL")}
        os.environ[env] = v
        if env == "ZPQ_DEC_VARIANT":
            os.environ["ZPQ_DEC_KERNEL"] = "6"          # the variant bits belong to the round-1 two-warp decoder
        try:
            ctx = Z.Context(0)
        finally:
            for k, o in old.items():
                if o is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = o
        try:
            for method in ("4", "5", CFG_ALL_METHOD):
                hdrs = {}
                for p in payloads:
                    hdrs.setdefault(Z.model_from_method(method, p), []).append(p)
                for hdr, ps in hdrs.items():
                    _roundtrip(ctx, hdr, ps)
        finally:
            ctx.close()


@pytest.mark.timeout(1500)
def test_full_size_mixed_models_in_one_launch(ctx):
    """BASELINE-size blocks (1 044 480 B), >= 256 of them, FOUR different models (three step signatures) in ONE batch: level 5
    on text (27 components: the period model fires), on 'mixed' and on random bytes (23), level 4 on text (8) and the
    17-component chain.
    One block of every model is compared with the oracle byte for byte; every block must decode back to its input."""
    n = (1 << 20) - 4096
    plan = [("5", 0, 64), ("5", 1, 64), ("5", 2, 48), ("4", 0, 48), (CFG_ALL_METHOD, 1, 40)]
    hdrs, datas, first = [], [], []
    for method, kind, count in plan:
        first.append(len(datas))
        for i in range(count):
            p = O.gen_block(kind, 9000 + 97 * kind + i, n)
            hdrs.append(Z.model_from_method(method, p))
            datas.append(b"\0" + p)
    assert len(datas) >= 256 and len(set(hdrs)) >= 4
    assert hdrs[first[0]][6] == 27 and hdrs[first[1]][6] == 23
    enc = ctx.encode_blocks(hdrs, datas)
    assert all(st == 0 for st, _ in enc)
    assert ctx.tiers()["wave_enc"] == len(datas)
    for k in first:
        assert enc[k][1][0] == O.encode(hdrs[k], datas[k]), "block %d differs from the oracle" % k
    dec = ctx.decode_blocks(hdrs, [segs[0] + b"\0\0\0\0" for _, segs in enc], caps=[len(d) + 16 for d in datas])
    assert ctx.tiers()["dec7"] == len(datas)
    assert all(st == 0 and segs[0] == datas[i] for i, (st, segs) in enumerate(dec))
