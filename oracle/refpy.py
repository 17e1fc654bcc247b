"""oracle/refpy.py -- ctypes binding of oracle/_ref/libzpaq_ref.so (TEST INFRASTRUCTURE).

The .so is the UNMODIFIED reference (libzpaq 7.15) compiled by oracle/Makefile
from /root/reference plus oracle/ref_driver.cpp.  Only tests/, bench.py's
reference / cpu_baseline legs and oracle/make_golden.py may import this.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def available(nojit=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libzpaq_ref_nojit.so" if nojit else "libzpaq_ref.so"))


def lib(nojit=False):
    global _lib
    if _lib is None:
        name = "libzpaq_ref_nojit.so" if (nojit or os.environ.get("ZPQ_REF_NOJIT")) else "libzpaq_ref.so"
        L = C.CDLL(os.path.join(_HERE, "_ref", name))
        L.ref_last_error.restype = C.c_char_p
        L.ref_compress_block.restype = C.c_longlong
        L.ref_compress_block.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_char_p, C.c_char_p,
                                         C.c_int, C.c_char_p, C.c_longlong]
        L.ref_compress_config.restype = C.c_longlong
        L.ref_compress_config.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.POINTER(C.c_int), C.c_int,
                                          C.c_int, C.c_int, C.c_char_p, C.c_longlong]
        L.ref_decompress.restype = C.c_longlong
        L.ref_decompress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong]
        L.ref_make_config.restype = C.c_longlong
        L.ref_make_config.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_longlong]
        L.ref_trace.restype = C.c_longlong
        L.ref_trace.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong, C.c_void_p]
        L.ref_bench.restype = C.c_double
        L.ref_bench.argtypes = [C.c_char_p, C.c_int, C.c_longlong, C.c_int, C.c_uint, C.c_int, C.c_int,
                                C.POINTER(C.c_longlong)]
        L.ref_bench_rt.restype = C.c_int
        L.ref_bench_rt.argtypes = [C.c_char_p, C.c_int, C.c_longlong, C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_double)]
        L.ref_gen_block.restype = None
        L.ref_gen_block.argtypes = [C.c_int, C.c_uint, C.c_char_p, C.c_longlong]
        _lib = L
    return _lib


def _err():
    return RuntimeError(lib().ref_last_error().decode("latin1"))


def compress_block(data, method, filename=b"", comment=None, dosha1=True):
    cap = len(data) + len(data) // 2 + 70000
    out = C.create_string_buffer(cap)
    n = lib().ref_compress_block(bytes(data), len(data), method.encode(), filename, comment, int(dosha1), out, cap)
    if n < 0:
        raise _err()
    return out.raw[:n]


def compress_config(data, config=None, args=None, level=0, nseg=1, dosha1=True):
    cap = len(data) + len(data) // 2 + 70000
    out = C.create_string_buffer(cap)
    a = (C.c_int * 9)(*(list(args or []) + [0] * 9)[:9])
    n = lib().ref_compress_config(bytes(data), len(data), (config or "").encode(), a, level, nseg, int(dosha1), out, cap)
    if n < 0:
        raise _err()
    return out.raw[:n]


def decompress(arc, cap):
    out = C.create_string_buffer(max(cap, 1))
    n = lib().ref_decompress(bytes(arc), len(arc), out, cap)
    if n < 0:
        raise _err()
    return out.raw[:n]


def make_config(method):
    out = C.create_string_buffer(1 << 20)
    a = (C.c_int * 9)()
    n = lib().ref_make_config(method.encode(), a, out, 1 << 20)
    if n < 0:
        raise _err()
    return out.value.decode(), list(a)


def trace(hdr, data):
    import numpy as np
    probs = np.zeros(len(data) * 8, dtype=np.uint16)
    n = lib().ref_trace(bytes(hdr), len(hdr), bytes(data), len(data), probs.ctypes.data)
    if n < 0:
        raise _err()
    return probs


def gen_block(kind, seed, n):
    out = C.create_string_buffer(max(n, 1))
    lib().ref_gen_block(kind, seed & 0xFFFFFFFF, out, n)
    return out.raw[:n]


def bench(method, nblocks, blocksize, kind, seed, threads, mode):
    tot = C.c_longlong(0)
    s = lib().ref_bench(method.encode(), nblocks, blocksize, kind, seed, threads, mode, C.byref(tot))
    if s < 0:
        raise _err()
    return s, tot.value


def bench_rt(method, nblocks, blocksize, kind, seed, threads):
    """(compress seconds, decompress seconds) of one pass over `nblocks` blocks on `threads` host threads."""
    t = (C.c_double * 2)()
    if lib().ref_bench_rt(method.encode(), nblocks, blocksize, kind, seed, threads, t) < 0:
        raise _err()
    return t[0], t[1]
