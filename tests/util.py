"""Shared helpers for the tests: golden fixtures, synthetic data, archive assembly."""
import base64
import hashlib
import json
import os

import oraclepy as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def data_of(spec):
    return O.gen_block(spec["kind"], spec["seed"], spec["n"])


def archive_bytes(entry):
    return base64.b64decode(entry["archive_b64"]) if "archive_b64" in entry else None


def assemble_block(hdr, coded_segments, payload_segments, filenames, comments, dosha1=True):
    """Block framing as Compressor writes it (SURVEY.md 8b block anatomy)."""
    out = bytearray(O.TAG + b"zPQ" + bytes([1 + (hdr[6] == 0), 1]) + hdr)
    for coded, pay, fn, cm in zip(coded_segments, payload_segments, filenames, comments):
        out += b"\x01" + fn + b"\0" + cm + b"\0\0" + coded + b"\0\0\0\0"
        out += (b"\xfd" + hashlib.sha1(pay).digest()) if dosha1 else b"\xfe"
    out += b"\xff"
    return bytes(out)


def split_segments(d, nseg):
    segs, pos = [], 0
    for s in range(nseg):
        ln = len(d) - pos if s == nseg - 1 else len(d) // nseg
        segs.append(d[pos:pos + ln])
        pos += ln
    return segs


class DevBytes:
    """Bytes in device memory for the *_dev entry points: a torch CUDA tensor on a GPU box; on the SIMT emulator (tools/simt,
    where "device" memory is host memory) a ctypes buffer."""

    def __init__(self, data=None, size=None):
        import ctypes as C
        n = len(data) if data is not None else size
        self.n = n
        try:
            import torch
            cuda = torch.cuda.is_available()
        except Exception:
            cuda = False
        if cuda:
            import torch
            if data is not None:
                self.t = torch.frombuffer(bytearray(bytes(data) + b"\0" * 16), dtype=torch.uint8).cuda()
            else:
                self.t = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            self.ptr = self.t.data_ptr()
        else:
            self.t = None
            self.buf = C.create_string_buffer(bytes(data) + b"\0" * 16 if data is not None else n + 16)
            self.ptr = C.cast(self.buf, C.c_void_p).value

    def bytes(self, n=None):
        n = self.n if n is None else n
        if self.t is not None:
            import torch
            torch.cuda.synchronize()
            return bytes(self.t[:n].cpu().numpy().tobytes())
        return self.buf.raw[:n]
