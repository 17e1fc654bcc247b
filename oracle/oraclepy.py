"""oracle/oraclepy.py -- ctypes binding of oracle/_build/libzpq_oracle.so plus a
tiny .zpaq block parser (TEST INFRASTRUCTURE).

The oracle is the CPU restatement of the reference hot path
(oracle/zpq_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product path never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzpq_oracle.so")
_lib = None

ERRORS = {0: "ok", -1: "bad header", -2: "bad component", -3: "ZPAQL execution error", -4: "output full",
          -5: "archive corrupted", -6: "unexpected end of file", -7: "out of memory", -8: "too big"}

TAG = bytes([0x37, 0x6B, 0x53, 0x74, 0xA0, 0x31, 0x83, 0xD3, 0x8C, 0xB2, 0x28, 0xB0, 0xD3])


def build():
    src = os.path.join(_HERE, "zpq_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
        L.zpqo_encode.argtypes = [vp, i, vp, sz, vp, sz, C.POINTER(sz)]
        L.zpqo_decode.argtypes = [vp, i, vp, sz, vp, sz, C.POINTER(sz)]
        L.zpqo_encode_segs.argtypes = [vp, i, vp, i, vp, vp, sz, vp]
        L.zpqo_decode_segs.argtypes = [vp, i, vp, i, vp, vp, sz, vp]
        L.zpqo_trace.argtypes = [vp, i, vp, sz, vp]
        L.zpqo_hcomp.argtypes = [vp, i, vp, sz, vp]
        L.zpqo_model_bytes.argtypes = [vp, i]
        L.zpqo_model_bytes.restype = C.c_double
        L.zpqo_tables.argtypes = [vp] * 5
        L.zpqo_table_sums.argtypes = [vp, vp]
        L.zpqo_gen_block.argtypes = [i, C.c_uint, vp, sz]
        for f in (L.zpqo_tables, L.zpqo_table_sums, L.zpqo_gen_block):
            f.restype = None
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, rc):
        super().__init__(ERRORS.get(rc, "error %d" % rc))
        self.rc = rc


def _buf(b):
    return C.cast(C.c_char_p(bytes(b)), C.c_void_p), bytes(b)


def encode(hdr, data, cap=None):
    """Coded bytes (incl. the 4 EOS flush bytes) for `data` (prefix included) under header `hdr`."""
    data = bytes(data)
    cap = cap or (len(data) + len(data) // 8 + 64)
    out = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    rc = lib().zpqo_encode(bytes(hdr), len(hdr), data, len(data), out, cap, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out.raw[:n.value]


def decode(hdr, coded, cap):
    out = C.create_string_buffer(max(cap, 1))
    n = C.c_size_t(0)
    rc = lib().zpqo_decode(bytes(hdr), len(hdr), bytes(coded), len(coded), out, cap, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out.raw[:n.value]


def encode_segs(hdr, segs):
    data = b"".join(segs)
    lens = (C.c_size_t * len(segs))(*[len(s) for s in segs])
    outl = (C.c_size_t * len(segs))()
    cap = len(data) + len(data) // 8 + 64 * len(segs)
    out = C.create_string_buffer(cap)
    rc = lib().zpqo_encode_segs(bytes(hdr), len(hdr), data, len(segs), lens, out, cap, outl)
    if rc:
        raise OracleError(rc)
    res, pos = [], 0
    for k in outl:
        res.append(out.raw[pos:pos + k])
        pos += k
    return res


def decode_segs(hdr, coded_segs, cap):
    data = b"".join(coded_segs)
    lens = (C.c_size_t * len(coded_segs))(*[len(s) for s in coded_segs])
    outl = (C.c_size_t * len(coded_segs))()
    out = C.create_string_buffer(max(cap, 1))
    rc = lib().zpqo_decode_segs(bytes(hdr), len(hdr), data, len(coded_segs), lens, out, cap, outl)
    if rc:
        raise OracleError(rc)
    res, pos = [], 0
    for k in outl:
        res.append(out.raw[pos:pos + k])
        pos += k
    return res


def trace(hdr, data):
    probs = np.zeros(len(data) * 8, dtype=np.uint16)
    rc = lib().zpqo_trace(bytes(hdr), len(hdr), bytes(data), len(data), probs.ctypes.data)
    if rc:
        raise OracleError(rc)
    return probs


def hcomp(hdr, data):
    n = bytes(hdr)[6]
    out = np.zeros((len(data), n), dtype=np.uint32)
    rc = lib().zpqo_hcomp(bytes(hdr), len(hdr), bytes(data), len(data), out.ctypes.data)
    if rc:
        raise OracleError(rc)
    return out


def model_bytes(hdr):
    return lib().zpqo_model_bytes(bytes(hdr), len(hdr))


def tables():
    ns = np.zeros(1024, np.uint8)
    dt2k = np.zeros(256, np.int32)
    dt = np.zeros(1024, np.int32)
    sq = np.zeros(4096, np.uint16)
    st = np.zeros(32768, np.int16)
    lib().zpqo_tables(ns.ctypes.data, dt2k.ctypes.data, dt.ctypes.data, sq.ctypes.data, st.ctypes.data)
    return dict(ns=ns, dt2k=dt2k, dt=dt, squash=sq, stretch=st)


def table_sums():
    a, b = C.c_uint(0), C.c_uint(0)
    lib().zpqo_table_sums(C.byref(a), C.byref(b))
    return a.value, b.value


def gen_block(kind, seed, n):
    out = C.create_string_buffer(max(n, 1))
    lib().zpqo_gen_block(kind, seed & 0xFFFFFFFF, out, n)
    return out.raw[:n]


KINDS = {"text": 0, "mixed": 1, "random": 2, "period": 3}


# ----------------------------------------------------------------------------
# Minimal .zpaq block parser (block anatomy: SURVEY.md section 8b).
class Segment:
    __slots__ = ("filename", "comment", "coded", "sha1")


class Block:
    __slots__ = ("level", "header", "segments", "end")


def parse_blocks(arc):
    """Split an archive into blocks -> [Block]; coded excludes the 4-zero terminator."""
    arc = bytes(arc)
    blocks, pos = [], 0
    while True:
        k = arc.find(TAG + b"zPQ", pos)
        if k < 0:
            break
        pos = k + 16
        b = Block()
        b.level = arc[pos]
        assert arc[pos + 1] == 1
        pos += 2
        hsize = arc[pos] + 256 * arc[pos + 1]
        b.header = arc[pos:pos + hsize + 2]
        pos += hsize + 2
        b.segments = []
        while arc[pos] == 1:
            s = Segment()
            pos += 1
            e = arc.index(b"\0", pos)
            s.filename = arc[pos:e]
            pos = e + 1
            e = arc.index(b"\0", pos)
            s.comment = arc[pos:e]
            pos = e + 1
            assert arc[pos] == 0
            pos += 1
            e = arc.index(b"\0\0\0\0", pos)
            while arc[e + 4] == 0:
                e += 1
            s.coded = arc[pos:e]
            pos = e + 4
            if arc[pos] == 253:
                s.sha1 = arc[pos + 1:pos + 21]
                pos += 21
            else:
                assert arc[pos] == 254
                s.sha1 = None
                pos += 1
            b.segments.append(s)
        assert arc[pos] == 255
        pos += 1
        b.end = pos
        blocks.append(b)
    return blocks
