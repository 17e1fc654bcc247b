"""GPU: randomly composed makeConfig methods (context models, ISSE chains, word models, MATCH, mixers, MIX2, SSE) against
archives written by the unmodified reference in the same run (oracle/_ref travels to the GPU box)."""
import random

import pytest

import oraclepy as O
import zpaq_b200 as Z

pytestmark = pytest.mark.gpu


def rand_method(rnd):
    parts = ["x0,0"]
    for _ in range(rnd.randint(1, 4)):
        k = rnd.choice(["ci", "c", "cm", "w", "a"])
        if k == "ci":
            parts.append("ci" + ",".join(str(rnd.choice([1, 1, 1, 2])) for _ in range(rnd.randint(1, 5))))
        elif k == "c":
            parts.append("c%d,%d,%d" % (rnd.choice([0, 0, 256, 4, 16]), rnd.choice([0, 0, 255, 1000 + rnd.randint(1, 40)]), rnd.choice([255, 255, 15, 0])))
            if rnd.random() < 0.5:
                parts.append("i%d" % rnd.choice([1, 2]))
        elif k == "cm":
            parts.append("c%d" % rnd.choice([0, 8, 256]))
        elif k == "w":
            parts.append("w%d" % rnd.choice([1, 2]))
            if rnd.random() < 0.5:
                parts.append("i1")
        else:
            parts.append("a")
    parts.append(rnd.choice(["m", "m16,24", "mm16", "m8t", "m16ts19", "mm16ts19t0", "t", "ms", "m0,24"]))
    return "".join(parts)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_methods_match_the_reference(ctx, seed):
    refpy = pytest.importorskip("refpy")
    if not refpy.available():
        pytest.skip("oracle/_ref not built")
    rnd = random.Random(seed)
    checked = 0
    for _ in range(12):
        method = rand_method(rnd)
        pays = [O.gen_block(rnd.randint(0, 3), rnd.randint(1, 9999), rnd.choice([0, 1, 50, 900, 2500, 4200])) for _ in range(3)]
        try:
            want = [refpy.compress_block(p, method) for p in pays]
        except Exception:
            continue                      # the reference itself rejects the string
        got = ctx.compress_blocks(method, pays)
        assert got == want, method
        out, nb = ctx.decompress_archive(b"".join(want), sum(map(len, pays)) + 16)
        assert nb == len(pays) and out == b"".join(pays), method
        checked += 1
    assert checked >= 6
