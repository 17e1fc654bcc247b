#!/usr/bin/env python
"""bench.py -- compress+decompress throughput of the B200 context-mixing core.

Contract (see the task statement): one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # libzpaq 7.15 on the host cores

Workload (BASELINE.json configs[2]): 1 GiB of synthetic 1 MiB-class blocks
(payload 1 044 480 B = 2^20 - 4096, what `zpaq -method 50` produces) coded with
-method 5, per GPU (weak scaling: blocks are independent, every rank codes its
own 1 GiB, no data-path collective).  A "step" is one full compress pass plus
one full decompress pass over that batch through the public C ABI
(zpq_compress_blocks / zpq_decompress_archive) with HOST buffers:
  * `e2e`   = payload bytes / wall time of those two calls (H2D/D2H inside),
  * `value` = the same bytes / device time of the kernels alone (inputs already
              in HBM when the kernels start), from CUDA events on the library's
              stream (zpq_last_timing),
  * `roofline` = algorithmic HBM bytes of the coding kernels / their duration.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

BLOCK = (1 << 20) - 4096
KIND = 1            # "mixed" generator
SEED = 777
METHOD = "5"
METRIC = "compress+decompress MB/s (-method 5), 1 GiB of 1 MiB blocks per GPU, bit-exact vs CPU ref"


def alg_bytes_per_input_byte(hdr, block_bytes):
    """B_alg = 8*sum_bit + 2*sum_nibble + sum_byte + model_bytes/block_bytes."""
    clen = [0, 2, 3, 2, 3, 4, 6, 6, 3, 5]
    n = hdr[6]
    pos = 7
    per_bit = per_nib = per_byte = 0
    mem = 2 ** (hdr[2] + 2) + 2 ** hdr[3]
    for _ in range(n):
        t, a = hdr[pos], hdr[pos + 1:pos + 6]
        size = 2 ** a[0]
        if t == 2:
            per_bit += 8; mem += 4 * size
        elif t == 3:
            per_bit += 10; per_nib += 16; mem += 64 * size + 1024
        elif t == 4:
            per_bit += 3; per_byte += 8; mem += 4 * size + 2 ** a[1]
        elif t == 6:
            per_bit += 4; mem += 2 * size
        elif t == 7:
            per_bit += 8 * a[2]; mem += 4 * size * a[2]
        elif t == 8:
            per_bit += 18; per_nib += 16; mem += 64 * size + 2048
        elif t == 9:
            per_bit += 12; mem += 128 * size
        pos += clen[t]
    per_byte += 2   # stream in + (about) one coded byte out
    return 8 * per_bit + 2 * per_nib + per_byte + mem / block_bytes


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled once a second DURING the timed region.  In-process NVML
    (nvidia_ml_py): a query is microseconds and takes no driver-wide lock for long; spawning nvidia-smi twice a second per
    rank initialises a driver context each time and measurably disturbs the launch path of the other ranks (the encoder
    issues ~525 launches per step).  nvidia-smi is only the fallback when NVML cannot be imported."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self.how = None
        self._stop_evt = threading.Event()

    def _nvml(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                idx = int(ids[self.index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        self.how = "nvml"
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                for nm, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(1.0)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.how = "nvidia-smi"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for (nm, _), v in zip(self.REASONS, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(2.0)

    def run(self):
        try:
            self._nvml()
        except Exception:
            if not self.samples:
                self._smi()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s), "source": self.how}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_lib():
    import refpy
    if not refpy.available():
        return None
    try:
        L = refpy.lib()
        refpy.compress_block(b"probe" * 100, "4")      # the x86 JIT needs RWX pages; fall back to the interpreter build
        return refpy
    except Exception:
        os.environ["ZPQ_REF_NOJIT"] = "1"
        refpy._lib = None
        return refpy


def cpu_reference_sample(threads, blocks_per_thread=1, block_bytes=BLOCK):
    """libzpaq 7.15 compress+decompress of `threads*blocks_per_thread` blocks, one block per thread at a time.
    `block_bytes` < BLOCK shortens every block (same generator, same method) to bound the sample's run time."""
    refpy = reference_lib()
    nblk = max(1, threads * blocks_per_thread)
    if refpy is not None:
        t_c, _ = refpy.bench(METHOD + "0", nblk, block_bytes, KIND, SEED, threads, 0)
        t_all, _ = refpy.bench(METHOD + "0", nblk, block_bytes, KIND, SEED, threads, 1)
        t_d = t_all  # mode 1 times the decompress phase only (archives prepared untimed)
        kind = "reference"
    else:
        import oraclepy as O
        import zpaq_b200 as Z
        t0 = time.time()
        data = O.gen_block(KIND, SEED, BLOCK)
        hdr = Z.model_from_method(METHOD, data)
        t0 = time.time(); coded = O.encode(hdr, b"\0" + data); t_c = time.time() - t0
        t0 = time.time(); O.decode(hdr, coded + b"\0\0\0\0", BLOCK + 1); t_d = time.time() - t0
        nblk, threads, kind, block_bytes = 1, 1, "port", BLOCK
    mb = nblk * block_bytes / 1e6
    return {"value": mb / (t_c + t_d), "unit": "MB/s", "cores": threads, "kind": kind,
            "compress_MBps": mb / t_c, "decompress_MBps": mb / t_d,
            "sample": "%d blocks x %d B, method %s, '%s' data, %d host threads (one block per thread), compress then decompress"
                      % (nblk, block_bytes, METHOD, "mixed", threads)}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = host_threads()
    # The CPU coder has no warm state (every block starts from freshly initialised tables), so the W warm-up steps are not
    # run.  A step = one block per host thread, compressed then decompressed; one full-size step is about a minute of CPU
    # time, so with more than three steps every block is shortened in proportion (same generator, same method): the whole
    # arm stays within a few minutes and MB/s per thread barely depends on the block length.
    block_bytes = BLOCK if args.steps <= 3 else max(1 << 16, (BLOCK * 3 // args.steps) & ~4095)
    best = []
    t_begin = time.time()
    for i in range(args.steps):
        best.append(cpu_reference_sample(threads, block_bytes=block_bytes))
    dt = time.time() - t_begin
    v = sum(b["value"] for b in best) / len(best)
    cb = dict(best[-1]); cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "-method 5 (makeConfig level-5 CM chain), 1 MiB-class blocks, bounded sample per step (see cpu_baseline.sample)",
                       "block_bytes": BLOCK, "method": METHOD},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--blocks", type=int, default=0, help="blocks per GPU (default: 1 GiB worth = 1029)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import oraclepy as O
    import zpaq_b200 as Z

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    total = 1 << 30
    nblk = args.blocks or (total + BLOCK - 1) // BLOCK
    sizes = [BLOCK] * nblk
    if not args.blocks:
        sizes[-1] = total - BLOCK * (nblk - 1)
    nbytes = sum(sizes)

    # ---- synthetic input in PINNED host memory (torch owns it) ----
    t0 = time.time()
    host_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    base = host_in.data_ptr()
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Lo = O.lib()
    for i in range(nblk):
        Lo.zpqo_gen_block(KIND, (SEED + rank * 100003 + i) & 0xFFFFFFFF, C.c_void_p(base + int(offs[i])), sizes[i])
    bound = int(Z.lib().zpq_compress_bound(BLOCK)) + 4096
    host_arc = torch.empty(nblk * bound, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    gen_s = time.time() - t0

    ctx = Z.Context(local_rank)
    L = Z.lib()
    dptr = (C.c_void_p * nblk)(*[base + int(offs[i]) for i in range(nblk)])
    lens = (C.c_uint64 * nblk)(*sizes)
    optr = (C.c_void_p * nblk)(*[host_arc.data_ptr() + i * bound for i in range(nblk)])
    ocap = (C.c_uint64 * nblk)(*([bound] * nblk))
    olen = (C.c_uint64 * nblk)()
    stat = (C.c_int32 * nblk)()
    arc_cat = torch.empty(nblk * bound, dtype=torch.uint8, pin_memory=True)

    def one_step():
        """compress every block, concatenate the archive, decompress it; returns timing dict."""
        t0 = time.perf_counter()
        rc = L.zpq_compress_blocks(ctx._h, nblk, METHOD.encode(), dptr, lens, None, None, 1, optr, ocap, olen, stat)
        t1 = time.perf_counter()
        if rc:
            raise SystemExit("zpq_compress_blocks failed: %d %s" % (rc, ctx.last_error()))
        tc = ctx.timing()
        # archive = blocks back to back (what zpaq's writer thread appends, zpaq.cpp:1968)
        pos = 0
        src = host_arc.numpy(); dst = arc_cat.numpy()
        for i in range(nblk):
            n = olen[i]
            dst[pos:pos + n] = src[i * bound:i * bound + n]
            pos += n
        n_out = C.c_uint64(0); nb = C.c_int(0)
        t2 = time.perf_counter()
        rc = L.zpq_decompress_archive(ctx._h, C.c_void_p(arc_cat.data_ptr()), pos, C.c_void_p(host_out.data_ptr()), nbytes,
                                      C.byref(n_out), C.byref(nb), 1)
        t3 = time.perf_counter()
        if rc:
            raise SystemExit("zpq_decompress_archive failed: %d %s" % (rc, ctx.last_error()))
        td = ctx.timing()
        if n_out.value != nbytes or nb.value != nblk:
            raise SystemExit("round trip size mismatch")
        return {"wall_c": t1 - t0, "wall_d": t3 - t2, "dev_c": (tc["init_ms"] + tc["code_ms"]) / 1e3,
                "dev_d": (td["init_ms"] + td["code_ms"]) / 1e3, "code_c": tc["code_ms"] / 1e3, "code_d": td["code_ms"] / 1e3,
                "arc_bytes": pos}

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = one_step()
    # round trip must be exact
    if not bool(torch.equal(host_in, host_out)) and args.warmup > 0:
        raise SystemExit("round trip mismatch")

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    barrier()
    acc = {"wall_c": 0.0, "wall_d": 0.0, "dev_c": 0.0, "dev_d": 0.0, "code_c": 0.0, "code_d": 0.0}
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        r = one_step()
        for k in acc:
            acc[k] += r[k]
    barrier()
    wall = time.perf_counter() - t_begin
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    if not bool(torch.equal(host_in, host_out)):
        raise SystemExit("round trip mismatch")

    # max over ranks of every time; totals over ranks of bytes
    times = torch.tensor([acc["wall_c"], acc["wall_d"], acc["dev_c"], acc["dev_d"], acc["code_c"], acc["code_d"], wall],
                         dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall_c, wall_d, dev_c, dev_d, code_c, code_d, wall = [float(x) for x in times.cpu()]
    tot_bytes = nbytes * world * args.steps
    mb = tot_bytes / 1e6

    if rank == 0:
        hdr = Z.model_from_method(METHOD, bytes(host_in.numpy()[:sizes[0]]))
        balg = alg_bytes_per_input_byte(hdr, BLOCK)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernel = the coding kernels (one compress + one decompress launch per step, per rank)
        per_rank_alg = nbytes * args.steps * balg * 2          # both directions touch the same tables
        achieved = per_rank_alg / (code_c + code_d) / 1e9
        # bit-exactness against the CPU reference on a bounded sample, and the CPU baseline itself
        cpu = None
        bit_exact = None
        if not args.no_cpu_baseline:
            cpu = cpu_reference_sample(host_threads())
            refpy = reference_lib()
            if refpy is not None:
                k = min(4, nblk)
                ok = True
                for i in range(k):
                    want = refpy.compress_block(bytes(host_in.numpy()[offs[i]:offs[i + 1]]), METHOD)
                    got = bytes(host_arc.numpy()[i * bound:i * bound + olen[i]])
                    ok = ok and (want == got)
                bit_exact = ok
        line = {
            "metric": METRIC, "value": mb / (dev_c + dev_d), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: -method 5 (23-component makeConfig CM chain), %d blocks x %d B per GPU "
                                   "(1 GiB), 'mixed' LCG text, seeds %d+rank*100003+block" % (nblk, BLOCK, SEED),
                       "blocks_per_gpu": nblk, "block_bytes": BLOCK, "method": METHOD, "parallelism": "block-shard x%d, no data-path collective" % world,
                       "cache": "each step touches %.0f GB of model tables per GPU (>> 126 MB L2); tables are re-initialised every launch"
                                % (nblk * 88.07e6 / 1e9)},
            "compress_MBps": mb / dev_c, "decompress_MBps": mb / dev_d,
            "e2e": {"value": mb / (wall_c + wall_d), "unit": "MB/s", "compress_MBps": mb / wall_c, "decompress_MBps": mb / wall_d,
                    "h2d_bytes_per_step": int(nbytes + r["arc_bytes"]), "d2h_bytes_per_step": int(r["arc_bytes"] + nbytes)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "alg_bytes_per_input_byte": balg, "kernel": "zpq_wave_kernel (encode, 525 launches per step) + zpq_dec_kernel6 (decode)",
                         "kernel_variants": {"ZPQ_WAVE_VARIANT": os.environ.get("ZPQ_WAVE_VARIANT", "25 (default)"),
                                             "ZPQ_DEC_VARIANT": os.environ.get("ZPQ_DEC_VARIANT", "31 (default)")},
                         "note": "latency/issue bound, not bandwidth bound (see profiles/ and DESIGN.md section 6)"},
            "cpu_baseline": cpu, "bit_exact_vs_reference": bit_exact, "clocks": clocks,
            "archive_ratio": r["arc_bytes"] / nbytes, "setup_s": gen_s,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
