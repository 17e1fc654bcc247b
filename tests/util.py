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


def exe_like(seed, n):
    """Deterministic payload with x86-looking CALL / JMP patterns (E8 / E9 .. .. .. 00 / FF), including overlapping candidates
    and opcode runs, for the E8E9 transform (libzpaq.cpp:6450-6459)."""
    b = bytearray(O.gen_block(2, seed, n))
    for i in range(0, max(0, n - 8), 61):
        b[i] = 0xE8 + (i & 1)
        b[i + 4] = 0x00 if i % 3 else 0xFF
    for i in range(1000, min(n - 8, 1400), 2):          # candidates two apart: every rewrite changes the next test
        b[i] = 0xE8
    for i in range(2000, min(n - 8, 2300)):             # no three consecutive non-opcode positions: one long chain
        b[i] = 0xFF if i % 5 == 4 else 0xE9
    return bytes(b)


def transform_payload(spec):
    return exe_like(spec["seed"], spec["n"]) if spec["variant"] == "exe" else O.gen_block(spec["kind"], spec["seed"], spec["n"])
