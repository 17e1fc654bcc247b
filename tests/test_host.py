"""CPU: host-side logic of the product library -- C ABI surface, model selection (row C1), validation."""
import ctypes as C
import os
import re

import pytest

import oraclepy as O
import zpaq_b200 as Z
from util import load, data_of, ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "zpq.h")).read()
    names = set(re.findall(r"\b(zpq_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    L = Z.lib()
    for n in sorted(names):
        assert hasattr(L, n), "libzpq_b200.so does not export %s" % n
    assert L.zpq_abi_version() == 1


HEADERS = load("headers.json")


def test_model_from_method_matches_reference_headers():
    """compressBlock level 4/5 expansion + makeConfig + Compiler, byte for byte (168 cases)."""
    bad = []
    for e in HEADERS:
        d = O.gen_block(e["kind"], e["seed"], e["n"])
        if Z.model_from_method(e["method"], d).hex() != e["header"]:
            bad.append((e["method"], e["kind"], e["n"]))
    assert not bad, bad


@pytest.mark.parametrize("method", ["0", "1", "2", "3", "4,0,0", "5,10,2", "x0,3ci1", "s4", ""])
def test_methods_outside_the_context_model_family_are_refused(method):
    with pytest.raises(Z.ZpqError) as e:
        Z.model_from_method(method, b"abc" * 100)
    assert e.value.rc in (-13, -10)


def test_model_bytes_tracks_reference_memory_formula():
    for m in ("4", "5"):
        d = O.gen_block(1, 1, 5000)
        h = Z.model_from_method(m, d)
        mine, ref = Z.model_bytes(h), O.model_bytes(h)      # ZPAQL::memory() restated in the oracle
        assert ref <= mine <= ref + 64 * 1024                # arena = same tables + alignment padding


def _hdr(comp, hcomp=bytes([56]), hh=0, hm=0):
    n = len(comp)
    body = bytes([hh, hm, 0, 0, n]) + b"".join(bytes(c) for c in comp) + b"\0" + hcomp + b"\0"
    return bytes([len(body) & 255, len(body) >> 8]) + body


@pytest.mark.parametrize("comp,code", [
    ([[3, 27]], -2),                 # max size for ICM is 26
    ([[1, 128], [8, 8, 1]], -2),     # ISSE j >= i
    ([[1, 128], [7, 4, 0, 2, 24, 255]], -2),   # MIX m not in 1..i-j
    ([[1, 128], [9, 4, 0, 255, 1]], -2),       # SSE start > limit*4
    ([[5, 0, 0, 128]], -2),          # AVG j >= i
    ([[10, 1]], -1),                 # invalid component type
])
def test_invalid_components_are_rejected_like_predictor_init(comp, code):
    with pytest.raises(Z.ZpqError) as e:
        Z.model_bytes(_hdr(comp))
    assert e.value.rc == code


def test_no_gpu_means_error_not_fallback():
    """Without a usable device the library reports ZPQ_E_CUDA; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert Z.lib().zpq_create(0, C.byref(h)) == -9
    with pytest.raises(Z.ZpqError):
        Z.Context(0)


def test_kernel_plan_for_baseline_models():
    L = Z.lib()
    L.zpq_model_describe.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint32)]
    for m, ncomp in (("4", 8), ("5", 23)):
        h = Z.model_from_method(m, O.gen_block(1, 777, 20000))
        info = (C.c_uint32 * 8)()
        assert L.zpq_model_describe(h, len(h), info) == 0
        assert info[0] == ncomp
        assert info[3] == 1, "HCOMP of makeConfig models must run as lane-parallel recipes"
        assert info[5] == 1, "fast kernel plan expected"
        assert 38528 + 7 * info[2] <= 232448, "level-5 tables must fit 7 blocks per SM"


POST = load("postproc.json")


@pytest.mark.parametrize("entry", POST, ids=["%s_k%d" % (e["method"].split(",")[0] + str(i), e["data"]["kind"]) for i, e in enumerate(POST)])
def test_host_postprocessor_against_reference_archives(entry):
    """zpq_postprocess_block (PASS prefix / PCOMP programs: LZ77, BWT, E8E9 inverse transforms; PostProcessor::write,
    libzpaq.cpp:2193-2241) on archives written by the unmodified reference.  The model output comes from the CPU oracle
    (or, for unmodelled level-2 blocks, from unpacking the stored chunks), so no GPU is involved."""
    import base64
    arc = base64.b64decode(entry["archive_b64"])
    want = data_of(entry["data"])
    blk = O.parse_blocks(arc)[0]
    hdr = blk.header
    coded = blk.segments[0].coded
    if hdr[6] == 0:
        raw = arc[arc.index(coded):]                  # 4-byte big-endian length + bytes ... zero length (libzpaq.cpp:2138-2153)
        mo, pos = b"", 0
        while True:
            ln = int.from_bytes(raw[pos:pos + 4], "big"); pos += 4
            if ln == 0:
                break
            mo += raw[pos:pos + ln]; pos += ln
    else:
        mo = O.decode(hdr, coded + b"\0\0\0\0", 4 * len(want) + 70000)
    assert Z.postprocess_block(hdr, [mo]) == [want]


def test_postprocessor_rejects_absurd_pcomp_sizes():
    """ph / pm come straight from an untrusted header: values ZPAQL::init refuses (libzpaq.cpp:1009-1024, "H too big" /
    "M too big") must come back as a status, not as a shift by 255 or a terabyte allocation."""
    model_out = bytes([1, 1, 0, 56]) + b"abc"           # PROG, length 1, program = HALT, then data
    for ph, pm in ((255, 0), (0, 200), (33, 33)):
        hdr = bytes([8, 0, 9, 16, ph, pm, 0, 0, 56, 0])
        with pytest.raises(Z.ZpqError) as e:
            Z.postprocess_block(hdr, [model_out])
        assert e.value.rc == -2
    hdr = bytes([8, 0, 9, 16, 4, 8, 0, 0, 56, 0])
    assert Z.postprocess_block(hdr, [model_out]) == [b""]   # a program that never OUTs produces nothing
