"""zpaq_b200 -- Python view of libzpq_b200.so (the B200-native context-mixing core).

This package is a thin ctypes binding of the C ABI in include/zpq.h, used by the
tests and bench.py.  It contains no compute: every coding call goes through the
shared library into the CUDA kernels, and raises when the library (or a GPU) is
missing -- there is deliberately no Python/CPU fallback.

PyTorch is used by callers only for device memory, streams and
torch.distributed plumbing (see bench.py); nothing here needs it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzpq_b200.so")

ERR = {
    0: "ok", -1: "invalid block header", -2: "invalid component", -3: "ZPAQL execution error",
    -4: "output buffer too small", -5: "archive corrupted", -6: "unexpected end of file",
    -7: "Out of memory", -8: "unsupported", -9: "CUDA error", -10: "bad argument",
    -11: "decoding end of stream", -12: "checksum mismatch", -13: "Unsupported method",
}


class ZpqError(RuntimeError):
    def __init__(self, rc, detail=""):
        super().__init__("zpq: %s (%d)%s" % (ERR.get(rc, "?"), rc, (": " + detail) if detail else ""))
        self.rc = rc


class Segment(C.Structure):
    """zpq_segment (include/zpq.h): one segment of an archive as zpq_scan_archive reports it."""
    _fields_ = [("block", C.c_uint32), ("segment", C.c_uint32), ("hdr_off", C.c_uint64), ("hlen", C.c_uint32),
                ("components", C.c_uint32), ("coded_off", C.c_uint64), ("coded_len", C.c_uint64), ("size_hint", C.c_int64),
                ("has_sha1", C.c_int32), ("sha1", C.c_uint8 * 20)]


def scan_archive(arc):
    """[Segment] of an in-memory archive (host only, no device)."""
    arc = bytes(arc)
    n = C.c_int(0); nb = C.c_int(0)
    buf = C.create_string_buffer(arc, max(len(arc), 1))
    rc = lib().zpq_scan_archive(C.cast(buf, C.c_void_p), len(arc), None, 0, C.byref(n), C.byref(nb))
    if rc:
        raise ZpqError(rc)
    segs = (Segment * max(n.value, 1))()
    rc = lib().zpq_scan_archive(C.cast(buf, C.c_void_p), len(arc), C.cast(segs, C.c_void_p), n.value, C.byref(n), C.byref(nb))
    if rc:
        raise ZpqError(rc)
    return list(segs[:n.value])


class Job(C.Structure):
    _fields_ = [("hdr", C.c_void_p), ("hlen", C.c_int32), ("nseg", C.c_int32),
                ("data", C.c_void_p), ("len", C.c_uint64), ("seg_len", C.c_void_p),
                ("out", C.c_void_p), ("out_cap", C.c_uint64), ("seg_out", C.c_void_p),
                ("out_len", C.c_uint64), ("status", C.c_int32), ("reserved", C.c_int32)]


EXPORTS = ["zpq_abi_version", "zpq_create", "zpq_destroy", "zpq_strerror", "zpq_last_error",
           "zpq_set_workspace", "zpq_set_stream", "zpq_model_bytes", "zpq_model_from_method",
           "zpq_encode_blocks", "zpq_decode_blocks", "zpq_encode_blocks_dev", "zpq_decode_blocks_dev",
           "zpq_last_timing", "zpq_launch_count", "zpq_compress_blocks", "zpq_decompress_archive",
           "zpq_compress_bound", "zpq_postprocess_block", "zpq_model_describe", "zpq_last_tiers",
           "zpq_compress_blocks_dev", "zpq_analyze_dev", "zpq_scan_archive", "zpq_header_from_stats", "zpq_block_plan", "zpq_last_prep_timing", "zpq_last_transform_bytes"]

_lib = None


def lib():
    """Load libzpq_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libzpq_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C zpaq_b200/csrc)")
        L = C.CDLL(LIB_PATH)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.zpq_abi_version.restype = i32
        L.zpq_create.argtypes = [i32, C.POINTER(vp)]
        L.zpq_destroy.argtypes = [vp]
        L.zpq_destroy.restype = None
        L.zpq_strerror.argtypes = [i32]
        L.zpq_strerror.restype = C.c_char_p
        L.zpq_last_error.argtypes = [vp]
        L.zpq_last_error.restype = C.c_char_p
        L.zpq_set_workspace.argtypes = [vp, vp, u64]
        L.zpq_set_stream.argtypes = [vp, vp]
        L.zpq_model_bytes.argtypes = [C.c_char_p, i32, C.POINTER(u64)]
        L.zpq_model_from_method.argtypes = [C.c_char_p, vp, u64, C.c_char_p, i32, C.POINTER(i32)]
        for f in (L.zpq_encode_blocks, L.zpq_decode_blocks, L.zpq_encode_blocks_dev, L.zpq_decode_blocks_dev):
            f.argtypes = [vp, i32, C.POINTER(Job)]
        L.zpq_last_timing.argtypes = [vp, C.POINTER(C.c_float)]
        L.zpq_last_prep_timing.argtypes = [vp, C.POINTER(C.c_float)]
        L.zpq_last_transform_bytes.argtypes = [vp]
        L.zpq_last_transform_bytes.restype = u64
        L.zpq_launch_count.argtypes = [vp]
        L.zpq_last_tiers.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.zpq_launch_count.restype = u64
        L.zpq_compress_blocks.argtypes = [vp, i32, C.c_char_p, vp, vp, vp, vp, i32, vp, vp, vp, vp]
        L.zpq_compress_blocks_dev.argtypes = [vp, i32, C.c_char_p, vp, vp, vp, vp, i32, vp, vp, vp, vp]
        L.zpq_analyze_dev.argtypes = [vp, i32, vp, vp, vp, vp]
        L.zpq_scan_archive.argtypes = [vp, u64, vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.zpq_decompress_archive.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64), C.POINTER(i32), i32]
        L.zpq_block_plan.argtypes = [C.c_char_p, u64, vp, C.c_char_p, i32, C.POINTER(i32), C.POINTER(i32), C.c_char_p, i32, C.POINTER(i32)]
        L.zpq_compress_bound.argtypes = [u64]
        L.zpq_postprocess_block.argtypes = [C.c_char_p, i32, vp, vp, i32, vp, u64, C.POINTER(u64), vp]
        L.zpq_compress_bound.restype = u64
        _lib = L
    return _lib


def model_from_method(method, data=b""):
    """Header bytes the reference's compressBlock(method) would write for this payload."""
    data = bytes(data)
    out = C.create_string_buffer(70000)
    n = C.c_int(0)
    buf = C.create_string_buffer(data, len(data)) if data else None
    rc = lib().zpq_model_from_method(method.encode(), C.cast(buf, C.c_void_p) if buf else None, len(data),
                                     out, 70000, C.byref(n))
    if rc:
        raise ZpqError(rc)
    return out.raw[:n.value]


def block_plan(method, n, gaps=(0, 0)):
    """(header bytes, pre-processing kind 0 / 3 / 4 / 7, bytes that go through the model in front of the payload) of a
    method for a block of n bytes -- what zpq_compress_blocks builds per block."""
    hdr = C.create_string_buffer(70000)
    pre = C.create_string_buffer(70000)
    hl, pl, prep = C.c_int(0), C.c_int(0), C.c_int(0)
    g = (C.c_int32 * 2)(*gaps)
    rc = lib().zpq_block_plan(method.encode(), n, g, hdr, 70000, C.byref(hl), C.byref(prep), pre, 70000, C.byref(pl))
    if rc:
        raise ZpqError(rc)
    return hdr.raw[:hl.value], prep.value, pre.raw[:pl.value]


def postprocess_block(hdr, model_out_segments):
    """Host PostProcessor over the model output of one block (list of per-segment bytes) -> list of payload segments."""
    segs = [bytes(x) for x in model_out_segments]
    flat = b"".join(segs)
    a = C.create_string_buffer(flat, max(len(flat), 1))
    sl = (C.c_uint64 * len(segs))(*[len(x) for x in segs])
    so = (C.c_uint64 * len(segs))()
    n = C.c_uint64(0)
    cap = max(len(flat) * 4, 1 << 16)
    while True:
        out = C.create_string_buffer(cap)
        rc = lib().zpq_postprocess_block(bytes(hdr), len(hdr), _addr(a), C.cast(sl, C.c_void_p), len(segs), _addr(out), cap, C.byref(n), C.cast(so, C.c_void_p))
        if rc == -4 and n.value > cap:
            cap = n.value
            continue
        if rc:
            raise ZpqError(rc)
        break
    res, pos = [], 0
    for k in range(len(segs)):
        res.append(out.raw[pos:pos + so[k]])
        pos += so[k]
    return res


def model_bytes(hdr):
    n = C.c_uint64(0)
    rc = lib().zpq_model_bytes(bytes(hdr), len(hdr), C.byref(n))
    if rc:
        raise ZpqError(rc)
    return n.value


def _addr(b):
    return C.cast(b, C.c_void_p).value


class Context:
    """One GPU context (zpq_ctx).  Not thread safe; one per device per process."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().zpq_create(device, C.byref(self._h))
        if rc:
            raise ZpqError(rc, "no usable CUDA device; this library has no CPU fallback")
        self.device = device

    def close(self):
        if self._h:
            lib().zpq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self):
        return lib().zpq_last_error(self._h).decode("latin1")

    def timing(self):
        ms = (C.c_float * 3)()
        lib().zpq_last_timing(self._h, ms)
        return {"init_ms": ms[0], "code_ms": ms[1], "call_ms": ms[2]}

    def prep_timing(self):
        """Device time of the stages in front of the coder in the last compress_blocks call: SHA-1 + statistics, E8E9 / BWT."""
        ms = (C.c_float * 2)()
        lib().zpq_last_prep_timing(self._h, ms)
        return {"analysis_ms": ms[0], "transform_ms": ms[1], "transform_alg_bytes": int(lib().zpq_last_transform_bytes(self._h))}

    def tiers(self):
        """Blocks of the last batch per kernel tier (zpq_last_tiers): wave encoder, dec7, dec6, fast, generic, wide generic."""
        t = (C.c_uint32 * 7)()
        lib().zpq_last_tiers(self._h, t)
        return dict(zip(("wave_enc", "dec7", "dec6", "fast", "generic", "wide", "dec8"), [int(x) for x in t]))

    def launch_count(self):
        return int(lib().zpq_launch_count(self._h))

    def set_workspace(self, ptr, nbytes):
        rc = lib().zpq_set_workspace(self._h, ptr, nbytes)
        if rc:
            raise ZpqError(rc)

    def set_stream(self, stream_ptr):
        rc = lib().zpq_set_stream(self._h, stream_ptr)
        if rc:
            raise ZpqError(rc)

    # -- hot path, host buffers ------------------------------------------------
    def _run(self, fn, hdrs, datas, seg_lens, caps):
        n = len(datas)
        jobs = (Job * n)()
        keep = []
        for i in range(n):
            h = C.create_string_buffer(bytes(hdrs[i]), len(hdrs[i]))
            d = C.create_string_buffer(bytes(datas[i]), max(len(datas[i]), 1))
            o = C.create_string_buffer(max(caps[i], 1))
            sl = seg_lens[i] if seg_lens and seg_lens[i] else [len(datas[i])]
            sa = (C.c_uint64 * len(sl))(*sl)
            so = (C.c_uint64 * len(sl))()
            keep.append((h, d, o, sa, so))
            jobs[i].hdr = _addr(h)
            jobs[i].hlen = len(hdrs[i])
            jobs[i].nseg = len(sl)
            jobs[i].data = _addr(d)
            jobs[i].len = len(datas[i])
            jobs[i].seg_len = _addr(sa)
            jobs[i].out = _addr(o)
            jobs[i].out_cap = caps[i]
            jobs[i].seg_out = _addr(so)
        rc = fn(self._h, n, jobs)
        if rc:
            raise ZpqError(rc, self.last_error())
        res = []
        for i in range(n):
            h, d, o, sa, so = keep[i]
            segs, pos = [], 0
            if jobs[i].status == 0:
                for k in so:
                    segs.append(o.raw[pos:pos + k])
                    pos += k
            res.append((jobs[i].status, segs))
        return res

    def encode_blocks(self, hdrs, datas, seg_lens=None, caps=None):
        """[(status, [coded segment bytes])] for model-input byte strings (prefix included)."""
        caps = caps or [len(d) + len(d) // 8 + 1024 for d in datas]
        return self._run(lib().zpq_encode_blocks, hdrs, datas, seg_lens, caps)

    def decode_blocks(self, hdrs, datas, seg_lens=None, caps=None):
        """datas: coded segments each followed by 4 zero bytes, concatenated per block."""
        return self._run(lib().zpq_decode_blocks, hdrs, datas, seg_lens, caps)

    # -- whole blocks ----------------------------------------------------------
    def compress_blocks(self, method, payloads, filenames=None, comments=None, dosha1=True):
        """Batched libzpaq::compressBlock -> list of complete .zpaq block byte strings."""
        n = len(payloads)
        bufs = [C.create_string_buffer(bytes(p), max(len(p), 1)) for p in payloads]
        caps = [int(lib().zpq_compress_bound(len(p))) + 70000 for p in payloads]
        outs = [C.create_string_buffer(c) for c in caps]
        dptr = (C.c_void_p * n)(*[_addr(b) for b in bufs])
        lens = (C.c_uint64 * n)(*[len(p) for p in payloads])
        optr = (C.c_void_p * n)(*[_addr(b) for b in outs])
        ocap = (C.c_uint64 * n)(*caps)
        olen = (C.c_uint64 * n)()
        st = (C.c_int32 * n)()
        fn = (C.c_char_p * n)(*[(f if f is not None else None) for f in filenames]) if filenames else None
        cm = (C.c_char_p * n)(*[(c if c is not None else None) for c in comments]) if comments else None
        rc = lib().zpq_compress_blocks(self._h, n, method.encode(), dptr, lens, fn, cm, int(dosha1), optr, ocap, olen, st)
        if rc:
            raise ZpqError(rc, self.last_error())
        return [outs[i].raw[:olen[i]] for i in range(n)]

    # -- device-resident (callers pass raw device pointers, e.g. torch tensors' data_ptr()) ------------------------------
    def analyze_dev(self, ptrs, lens, sha1=True, gaps=True):
        """(list of 20-byte digests or None, list of (gap0, gap1) or None) for payloads in HBM."""
        n = len(ptrs)
        pa = (C.c_void_p * n)(*ptrs)
        la = (C.c_uint64 * n)(*lens)
        sb = (C.c_uint8 * (20 * n))() if sha1 else None
        gb = (C.c_int32 * (2 * n))() if gaps else None
        rc = lib().zpq_analyze_dev(self._h, n, C.cast(pa, C.c_void_p), C.cast(la, C.c_void_p),
                                   C.cast(sb, C.c_void_p) if sha1 else None, C.cast(gb, C.c_void_p) if gaps else None)
        if rc:
            raise ZpqError(rc, self.last_error())
        return ([bytes(sb[20 * i:20 * i + 20]) for i in range(n)] if sha1 else None,
                [(gb[2 * i], gb[2 * i + 1]) for i in range(n)] if gaps else None)

    def decode_blocks_dev(self, hdrs, ptrs, lens, out_ptrs, out_caps, seg_lens=None):
        """zpq_decode_blocks_dev: coded segments (each with its terminator zeros) and outputs in HBM.
        Returns [(status, out_len, [segment output lengths])]."""
        n = len(ptrs)
        jobs = (Job * n)()
        keep = []
        for i in range(n):
            h = C.create_string_buffer(bytes(hdrs[i]), len(hdrs[i]))
            sl = seg_lens[i] if seg_lens and seg_lens[i] else [lens[i]]
            sa = (C.c_uint64 * len(sl))(*sl)
            so = (C.c_uint64 * len(sl))()
            keep.append((h, sa, so))
            jobs[i].hdr = _addr(h); jobs[i].hlen = len(hdrs[i]); jobs[i].nseg = len(sl)
            jobs[i].data = ptrs[i]; jobs[i].len = lens[i]; jobs[i].seg_len = _addr(sa)
            jobs[i].out = out_ptrs[i]; jobs[i].out_cap = out_caps[i]; jobs[i].seg_out = _addr(so)
        rc = lib().zpq_decode_blocks_dev(self._h, n, jobs)
        if rc:
            raise ZpqError(rc, self.last_error())
        return [(jobs[i].status, int(jobs[i].out_len), list(keep[i][2])) for i in range(n)]

    def compress_blocks_dev(self, method, ptrs, lens, out_ptrs, out_caps, dosha1=True):
        """zpq_compress_blocks_dev: payloads and output blocks in HBM; returns (status list, block length list)."""
        n = len(ptrs)
        pa = (C.c_void_p * n)(*ptrs)
        la = (C.c_uint64 * n)(*lens)
        oa = (C.c_void_p * n)(*out_ptrs)
        ca = (C.c_uint64 * n)(*out_caps)
        ol = (C.c_uint64 * n)()
        st = (C.c_int32 * n)()
        rc = lib().zpq_compress_blocks_dev(self._h, n, method.encode(), C.cast(pa, C.c_void_p), C.cast(la, C.c_void_p), None, None,
                                           int(dosha1), C.cast(oa, C.c_void_p), C.cast(ca, C.c_void_p), C.cast(ol, C.c_void_p), C.cast(st, C.c_void_p))
        if rc:
            raise ZpqError(rc, self.last_error())
        return list(st), list(ol)

    def decompress_archive(self, arc, cap, verify=True):
        arc = bytes(arc)
        a = C.create_string_buffer(arc, len(arc))
        out = C.create_string_buffer(max(cap, 1))
        n = C.c_uint64(0)
        nb = C.c_int(0)
        rc = lib().zpq_decompress_archive(self._h, _addr(a), len(arc), _addr(out), cap, C.byref(n), C.byref(nb), int(verify))
        if rc:
            raise ZpqError(rc, self.last_error())
        return out.raw[:n.value], nb.value
