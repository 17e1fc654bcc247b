#!/usr/bin/env python
"""bench.py -- compress+decompress throughput of the B200 context-mixing core.

Contract (see the task statement): one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # libzpaq 7.15 on the host cores
  python bench.py --config m4_256 | chain17_512            # BASELINE configs[1] / configs[3] (not the headline line)
  python bench.py --config bwt_256                         # the BWT path, with the suffix sorter's own roofline

Workload (BASELINE.json configs[2], the default): 1 GiB of synthetic 1 MiB-class blocks
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

Run time: the whole run is kept inside ZPQ_BENCH_BUDGET_S (default 600 s) of wall clock.  The first warm-up step is a
full-size step (it allocates the workspace and brings the clocks up); the remaining W-1 warm-up steps run the same two
calls on all blocks cut to the first 1/16 of their length (same launches and grid; the kernels keep no state between
calls: every call re-initialises its tables; their duration follows the block length, not the block count).
If K timed steps would not fit the budget the run stops early and the line reports the steps actually timed
(`steps`) next to `steps_requested` -- a flagged line instead of a killed process.
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

T_START = time.time()
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

BLOCK = (1 << 20) - 4096
SEED = 777
CHAIN17 = "x0,0ci1,1,1,1,1,1,1,1c256c8,0,255c16,0,255,255aw1c0,0,255,1001,255c0,1010,255m16,24"
# name -> (method string for this library, method string for libzpaq::compressBlock, total bytes per GPU, generator, label)
CONFIGS = {
    "m5_1g": ("5", "50", 1 << 30, 1, "BASELINE configs[2]: -method 5 (23-component makeConfig CM chain), 1 GiB of 1 MiB-class blocks per GPU, 'mixed' LCG text"),
    "m4_256": ("4", "40", 256 * BLOCK, 0, "BASELINE configs[1]: -method 4 (8-component ICM/ISSE chain + MATCH + MIX), 256 x 1 MiB-class 'text' blocks"),
    "bwt_256": ("x0,3ci1", "x0,3ci1", 256 * BLOCK, 0, "SURVEY.md 8(f)-4 (not a BASELINE config): BWT pre-processing on the device (batched suffix sort, zpq_bwt.cu) + 2-component ICM/ISSE model, 256 x 1 MiB-class 'text' blocks; decompression inverts the BWT on the device"),
    "chain17_512": (CHAIN17, CHAIN17, 1 << 29, 1, "BASELINE configs[3]: custom 17-component chain (8-deep ISSE chain, 16-input MIX with 2^16 weight sets), 512 MiB of 1 MiB-class 'mixed' blocks"),
}
KIND_NAME = {0: "text", 1: "mixed", 2: "random"}
METRIC = "compress+decompress MB/s (-method 5), 1 GiB of 1 MiB blocks per GPU, bit-exact vs CPU ref"
BUDGET_S = float(os.environ.get("ZPQ_BENCH_BUDGET_S", "600"))


def metric_name(cfg):
    if cfg == "m5_1g":
        return METRIC
    return "compress+decompress MB/s (%s), bit-exact vs CPU ref" % cfg


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


def host_info():
    """What the host side of this lease really is: affinity, cgroup CPU quota, CPU model (a throttled lease and a whole host
    both show 128 in sched_getaffinity)."""
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = os.cpu_count() or 1
    quota = None
    raw = None
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            raw = open(path).read().strip()
            if path.endswith("cpu.max"):
                q, per = raw.split()
                if q != "max":
                    quota = float(q) / float(per)
            else:
                q = float(raw)
                per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    quota = q / per
            break
        except Exception:
            continue
    model = None
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    eff = aff if quota is None else max(1, min(aff, int(quota + 0.999)))
    par = measured_parallelism(aff)
    if par is not None:
        eff = max(1, min(eff, int(par + 0.75)))
    return {"affinity": aff, "nproc": os.cpu_count(), "cgroup_cpu_max": raw, "cgroup_quota_cpus": quota, "cpu_model": model,
            "measured_parallelism": par, "threads_used": eff}


_PAR = {}


def measured_parallelism(nthreads):
    """How many host cores this process really gets: a pure-compute loop (the block generator, 4 MB per thread) on
    `nthreads` threads against the same loop on one thread.  A lease with a CPU quota or shared cores shows the quota here
    even when the cgroup files are not visible.  ~0.1 s, cached."""
    if nthreads in _PAR:
        return _PAR[nthreads]
    try:
        import oraclepy as O
        Lo = O.lib()
        n = 2 << 20
        bufs = [C.create_string_buffer(n) for _ in range(nthreads)]

        def work(i):
            for rep in range(6):
                Lo.zpqo_gen_block(2, 12345 + i + rep, bufs[i], n)

        def timed(k):
            th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            return time.perf_counter() - t0
        timed(1)
        t1 = min(timed(1), timed(1))
        tn = min(timed(nthreads), timed(nthreads))
        _PAR[nthreads] = nthreads * t1 / tn
    except Exception:
        _PAR[nthreads] = None
    return _PAR[nthreads]


def host_threads():
    return host_info()["threads_used"]


def reference_lib():
    import refpy
    if not refpy.available():
        return None
    try:
        refpy.lib()
        refpy.compress_block(b"probe" * 100, "4")      # the x86 JIT needs RWX pages; fall back to the interpreter build
        return refpy
    except Exception:
        os.environ["ZPQ_REF_NOJIT"] = "1"
        refpy._lib = None
        return refpy


def cpu_reference_sample(cfg, threads, nblk=None):
    """libzpaq 7.15 compress+decompress of `nblk` FULL-SIZE blocks (default: one per thread) of the configuration's
    workload, one block per thread at a time as zpaq.cpp:1918 does; every block is compressed once and the archives just
    written are decompressed (one pass, ref_bench_rt)."""
    method, ref_method, _, kind, _ = CONFIGS[cfg]
    refpy = reference_lib()
    nblk = max(1, nblk or threads)
    info = host_info()
    if refpy is not None:
        t_c, t_d = refpy.bench_rt(ref_method, nblk, BLOCK, kind, SEED, threads)
        how = "reference"
    else:
        import oraclepy as O
        import zpaq_b200 as Z
        data = O.gen_block(kind, SEED, BLOCK)
        hdr = Z.model_from_method(method, data)
        t0 = time.time(); coded = O.encode(hdr, b"\0" + data); t_c = time.time() - t0
        t0 = time.time(); O.decode(hdr, coded + b"\0\0\0\0", BLOCK + 1); t_d = time.time() - t0
        nblk, threads, how = 1, 1, "port"
    mb = nblk * BLOCK / 1e6
    return {"value": mb / (t_c + t_d), "unit": "MB/s", "cores": threads, "kind": how,
            "compress_MBps": mb / t_c, "decompress_MBps": mb / t_d, "seconds": t_c + t_d, "host": info,
            "sample": "%d blocks x %d B (full size), %s, '%s' data, %d host threads (one block per thread at a time), compress then decompress"
                      % (nblk, BLOCK, "-method " + method if len(method) < 3 else "custom chain", KIND_NAME[kind], threads)}


def run_reference(args, rank):
    """Reference arm: the unmodified libzpaq on the box's host cores, same config / metric / unit.  A step = one full-size
    block per usable host thread, compressed and decompressed; the CPU coder keeps no state between blocks (every block
    starts from freshly initialised tables), so the W warm-up steps are one small untimed pass."""
    if rank != 0:
        return
    cfg = args.config
    method, ref_method, total, kind, label = CONFIGS[cfg]
    threads = host_threads()
    if args.warmup > 0:
        cpu_reference_sample(cfg, min(threads, 4), 4)
    t_begin = time.time()
    samples = []
    for i in range(args.steps):
        # bounded: stop when the next step would leave the budget (a step is ~10-30 s of wall clock)
        if samples and (time.time() - T_START) + 1.3 * samples[-1]["seconds"] > BUDGET_S * 0.6:
            break
        samples.append(cpu_reference_sample(cfg, threads))
    dt = time.time() - t_begin
    tot_mb = sum(s["seconds"] * s["value"] for s in samples)
    tot_s = sum(s["seconds"] for s in samples)
    v = tot_mb / tot_s
    cb = dict(samples[-1]); cb["value"] = v
    cb["compress_MBps"] = tot_mb / sum(s["seconds"] * s["value"] / s["compress_MBps"] for s in samples)
    cb["decompress_MBps"] = tot_mb / sum(s["seconds"] * s["value"] / s["decompress_MBps"] for s in samples)
    line = {"impl": "reference", "metric": metric_name(cfg), "value": v, "unit": "MB/s", "n_gpus": args.gpus, "steps": len(samples),
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(1, len(samples)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": label + "; reference arm: bounded sample per step = one full-size block per usable host thread (see cpu_baseline.sample)",
                       "block_bytes": BLOCK, "blocks_per_step": threads, "method": method},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_sharded(args, rank, local_rank, world, dist):
    """north_star's multi-GPU data path, measured: rank 0 holds every raw block in pinned host memory; per step they go up
    to its HBM, are scattered over NCCL, coded where they land (zpq_compress_blocks_dev), gathered and copied down; the
    archive then takes the same way for decompression.  One process per GPU; with --gpus 1 the collectives are no-ops."""
    import torch
    import zpaq_b200 as Z
    from zpaq_b200 import shard
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sharded_roundtrip import gen_blocks_tensor
    if dist is None:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    cfg = args.config
    METHOD, ref_method, total, KIND, label = CONFIGS[cfg]
    per_gpu = args.blocks or (total + BLOCK - 1) // BLOCK
    nblk = per_gpu * world
    dev = torch.device("cuda", local_rank)
    ctx = Z.Context(local_rank)

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    host = host_arc = host_out = None
    lens = None
    if rank == 0:
        host = gen_blocks_tensor(nblk, BLOCK, kind=KIND, seed=SEED)
        lens = [BLOCK] * nblk
        host_arc = torch.empty(nblk * (BLOCK // 2 + BLOCK // 8 + 80000), dtype=torch.uint8, pin_memory=True)
        host_out = torch.empty(nblk * BLOCK, dtype=torch.uint8, pin_memory=True)

    def step(nb):
        """one round trip over the first nb blocks; returns phase times"""
        t0 = time.perf_counter()
        flat = host[:nb * BLOCK].to(dev, non_blocking=True) if rank == 0 else None
        arc, alens, tc = shard.compress_sharded(ctx, METHOD, flat, lens[:nb] if rank == 0 else None, src=0, device=dev, sync=sync)
        if rank == 0:
            host_arc[:arc.numel()].copy_(arc, non_blocking=True)
        sync(); t1 = time.perf_counter()
        arc2 = host_arc[:arc.numel()].to(dev, non_blocking=True) if rank == 0 else None
        out, olens, td = shard.decompress_sharded(ctx, arc2, src=0, device=dev, sync=sync)
        if rank == 0:
            host_out[:out.numel()].copy_(out, non_blocking=True)
        sync(); t2 = time.perf_counter()
        return {"wall_c": t1 - t0, "wall_d": t2 - t1, "dev_c": tc["dev_s"], "dev_d": td["dev_s"], "scatter_c": tc["scatter_s"], "gather_c": tc["gather_s"],
                "scatter_d": td["scatter_s"], "gather_d": td["gather_s"], "arc": int(arc.numel()) if rank == 0 else 0}

    for w in range(args.warmup):
        r = step(nblk if w == 0 else max(world, nblk // 8))
    if rank == 0 and args.warmup and not bool(torch.equal(host, host_out)):
        raise SystemExit("sharded round trip mismatch")
    acc = {}
    done = 0
    sync(); t_begin = time.perf_counter()
    for k in range(args.steps):
        if done >= 1:
            over = torch.tensor([1.0 if (time.time() - T_START) + 1.2 * r["wall_c"] + 1.2 * r["wall_d"] + 30 > BUDGET_S else 0.0], device=dev)
            dist.all_reduce(over, op=dist.ReduceOp.MAX)
            if over.item() > 0:
                break
        r = step(nblk)
        done += 1
        for key, v in r.items():
            acc[key] = acc.get(key, 0.0) + v
    sync(); wall = time.perf_counter() - t_begin
    t = torch.tensor([acc["dev_c"], acc["dev_d"]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_c, dev_d = float(t[0]), float(t[1])
    if rank == 0:
        if not bool(torch.equal(host, host_out)):
            raise SystemExit("sharded round trip mismatch")
        mb = nblk * BLOCK * done / 1e6
        line = {"metric": metric_name(cfg), "mode": "sharded", "value": mb / (dev_c + dev_d), "unit": "MB/s", "n_gpus": world, "steps": done,
                "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, done), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": "%s; sharded data path: rank 0 holds %d blocks x %d B, NCCL scatter of raw blocks / gather of coded blocks"
                                       % (label, nblk, BLOCK), "blocks_per_gpu": per_gpu, "block_bytes": BLOCK, "method": METHOD,
                           "parallelism": "block-shard x%d; NCCL point-to-point scatter / gather around the coding path" % world},
                "compress_MBps": mb / dev_c, "decompress_MBps": mb / dev_d,
                "e2e": {"value": mb / (acc["wall_c"] + acc["wall_d"]), "unit": "MB/s", "compress_MBps": mb / acc["wall_c"], "decompress_MBps": mb / acc["wall_d"],
                        "h2d_bytes_per_step": int(nblk * BLOCK + acc["arc"] / done), "d2h_bytes_per_step": int(nblk * BLOCK + acc["arc"] / done)},
                "nccl": {"scatter_raw_s": acc["scatter_c"] / done, "gather_coded_s": acc["gather_c"] / done, "scatter_coded_s": acc["scatter_d"] / done,
                         "gather_raw_s": acc["gather_d"] / done, "scatter_raw_GBps": nblk * BLOCK / 1e9 / max(acc["scatter_c"] / done, 1e-9)},
                "total_s": time.time() - T_START}
        print(json.dumps(line), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="m5_1g", choices=sorted(CONFIGS))
    ap.add_argument("--blocks", type=int, default=0, help="blocks per GPU (default: the configuration's size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="local", choices=["local", "sharded"],
                    help="local (default, the contract line): every rank codes its own data; sharded: rank 0 holds all blocks, "
                         "NCCL scatters raw blocks / gathers coded blocks (north_star's data path), not the headline line")
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

    cfg = args.config
    METHOD, ref_method, total, KIND, label = CONFIGS[cfg]
    if args.mode == "sharded":
        run_sharded(args, rank, local_rank, world, dist)
        return
    nblk = args.blocks or (total + BLOCK - 1) // BLOCK
    sizes = [BLOCK] * nblk
    if not args.blocks and total % BLOCK:
        sizes[-1] = total - BLOCK * (nblk - 1)
    nbytes = sum(sizes)

    # ---- synthetic input in PINNED host memory (torch owns it) ----
    t0 = time.time()
    host_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    base = host_in.data_ptr()
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    Lo = O.lib()

    def gen(lo, hi):
        for i in range(lo, hi):
            Lo.zpqo_gen_block(KIND, (SEED + rank * 100003 + i) & 0xFFFFFFFF, C.c_void_p(base + int(offs[i])), sizes[i])
    gth = [threading.Thread(target=gen, args=(k * nblk // 8, (k + 1) * nblk // 8)) for k in range(8)]
    for t in gth:
        t.start()
    for t in gth:
        t.join()
    bound = int(Z.lib().zpq_compress_bound(BLOCK)) + 4096
    host_arc = torch.empty(nblk * bound, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    arc_cat = torch.empty(nblk * bound, dtype=torch.uint8, pin_memory=True)
    gen_s = time.time() - t0

    ctx = Z.Context(local_rank)
    L = Z.lib()
    dptr = (C.c_void_p * nblk)(*[base + int(offs[i]) for i in range(nblk)])
    lens = (C.c_uint64 * nblk)(*sizes)
    optr = (C.c_void_p * nblk)(*[host_arc.data_ptr() + i * bound for i in range(nblk)])
    ocap = (C.c_uint64 * nblk)(*([bound] * nblk))
    olen = (C.c_uint64 * nblk)()
    stat = (C.c_int32 * nblk)()

    short = max(4096, BLOCK // 16)
    lens_short = (C.c_uint64 * nblk)(*[min(short, n) for n in sizes])

    def one_step(nb=nblk, shortened=False):
        """compress the first `nb` blocks (`shortened`: only the first 1/16 of each), concatenate the archive,
        decompress it; returns timing dict."""
        nby = sum(lens_short[i] for i in range(nb)) if shortened else int(offs[nb])
        t0 = time.perf_counter()
        rc = L.zpq_compress_blocks(ctx._h, nb, METHOD.encode(), dptr, lens_short if shortened else lens, None, None, 1, optr, ocap, olen, stat)
        t1 = time.perf_counter()
        if rc:
            raise SystemExit("zpq_compress_blocks failed: %d %s" % (rc, ctx.last_error()))
        tc = ctx.timing()
        tp = ctx.prep_timing()
        # archive = blocks back to back (what zpaq's writer thread appends, zpaq.cpp:1968)
        pos = 0
        src = host_arc.numpy(); dst = arc_cat.numpy()
        for i in range(nb):
            n = olen[i]
            dst[pos:pos + n] = src[i * bound:i * bound + n]
            pos += n
        n_out = C.c_uint64(0); nbk = C.c_int(0)
        t2 = time.perf_counter()
        rc = L.zpq_decompress_archive(ctx._h, C.c_void_p(arc_cat.data_ptr()), pos, C.c_void_p(host_out.data_ptr()), nby,
                                      C.byref(n_out), C.byref(nbk), 1)
        t3 = time.perf_counter()
        if rc:
            raise SystemExit("zpq_decompress_archive failed: %d %s" % (rc, ctx.last_error()))
        td = ctx.timing()
        if n_out.value != nby or nbk.value != nb:
            raise SystemExit("round trip size mismatch")
        return {"wall_c": t1 - t0, "wall_d": t3 - t2, "dev_c": (tc["init_ms"] + tc["code_ms"]) / 1e3,
                "dev_d": (td["init_ms"] + td["code_ms"]) / 1e3, "code_c": tc["code_ms"] / 1e3, "code_d": td["code_ms"] / 1e3,
                "arc_bytes": pos, "wall": t3 - t0, "xf_s": tp["transform_ms"] / 1e3, "xf_bytes": float(tp["transform_alg_bytes"])}

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def agree(flag):
        """the same decision on every rank (max over ranks)"""
        if dist is None:
            return flag
        t = torch.tensor([1.0 if flag else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(t.item() > 0)

    # ---- warm-up: one full-size step (workspace allocation, clocks), then W-1 steps over ALL blocks cut to 1/16 of their
    # length (same launches, same grid, same workspace; the kernels' duration follows the block length, not the count) ----
    reserve_s = 30.0 + (0.0 if args.no_cpu_baseline else 20.0)       # CPU baseline, bit-exact sample, teardown, start-up slack
    step_est = None
    for w in range(args.warmup):
        full = (w == 0)
        r = one_step(nblk, shortened=not full)
        if full:
            step_est = r["wall"]
            if not bool(torch.equal(host_in, host_out)):
                raise SystemExit("round trip mismatch")
    warm_s = time.time() - T_START

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    barrier()
    acc = {"wall_c": 0.0, "wall_d": 0.0, "dev_c": 0.0, "dev_d": 0.0, "code_c": 0.0, "code_d": 0.0, "xf_s": 0.0, "xf_bytes": 0.0}
    t_begin = time.perf_counter()
    done = 0
    for k in range(args.steps):
        if step_est is not None and done >= 1:
            over = (time.time() - T_START) + 1.1 * step_est + reserve_s > BUDGET_S
            if agree(over):
                break
        r = one_step()
        step_est = r["wall"]
        done += 1
        for key in acc:
            acc[key] += r[key]
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
    tot_bytes = nbytes * world * done
    mb = tot_bytes / 1e6

    if rank == 0:
        try:
            hdr = Z.model_from_method(METHOD, bytes(host_in.numpy()[:sizes[0]]))
        except Z.ZpqError:
            hdr = Z.block_plan(METHOD, sizes[0])[0]          # a method that pre-processes the payload (BWT)
        balg = alg_bytes_per_input_byte(hdr, BLOCK)
        arena_bytes = Z.model_bytes(hdr)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernels = the coding kernels of the two directions (per rank)
        per_rank_alg = nbytes * done * balg * 2          # both directions touch the same tables
        achieved = per_rank_alg / (code_c + code_d) / 1e9
        # DRAM traffic of the two coding kernels from their ncu captures (bytes per input byte, profiles/r02_traffic.json),
        # scaled to what this rank coded: the same "per launch set" basis as `achieved`
        traffic = traffic_detail = None
        try:
            traffic_detail = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if cfg == "m5_1g":
                traffic = nbytes * done * (traffic_detail["encode_bytes_per_input_byte"] + traffic_detail["decode_bytes_per_input_byte"]) / (code_c + code_d) / 1e9
            else:
                traffic_detail = {"note": "ncu captures are of the -method 5 chain (profiles/r02_traffic.json); not measured for this configuration"}
        except Exception:
            pass
        # bit-exactness against the CPU reference on a bounded sample, and the CPU baseline itself
        cpu = None
        bit_exact = None
        if not args.no_cpu_baseline:
            refpy = reference_lib()
            if refpy is not None:
                k = min(4, nblk)
                want = [None] * k

                def ref_one(i):
                    want[i] = refpy.compress_block(bytes(host_in.numpy()[offs[i]:offs[i + 1]]), ref_method)
                th = [threading.Thread(target=ref_one, args=(i,)) for i in range(k)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                bit_exact = all(want[i] == bytes(host_arc.numpy()[i * bound:i * bound + olen[i]]) for i in range(k))
            left = BUDGET_S - (time.time() - T_START) - 15.0
            ht = host_threads()
            # one full-size block per usable host thread is ~10-30 s; with little budget left, fewer blocks
            cpu = cpu_reference_sample(cfg, ht, ht if left > 40 else max(1, ht // 4))
        line = {
            "metric": metric_name(cfg), "value": mb / (dev_c + dev_d), "unit": "MB/s", "n_gpus": world, "steps": done,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / max(1, done), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "%s; %d blocks x %d B per GPU, seeds %d+rank*100003+block" % (label, nblk, BLOCK, SEED),
                       "blocks_per_gpu": nblk, "block_bytes": BLOCK, "method": METHOD, "parallelism": "block-shard x%d, no data-path collective" % world,
                       "cache": "inputs larger than L2: each step touches %.1f GB of model tables per GPU (>> 126 MB L2); tables are re-initialised every launch"
                                % (nblk * arena_bytes / 1e9),
                       "warmup_detail": "step 1 full size, steps 2..W on all blocks cut to their first %d bytes" % short,
                       "budget_s": BUDGET_S, "truncated": done < args.steps},
            "compress_MBps": mb / dev_c, "decompress_MBps": mb / dev_d,
            "e2e": {"value": mb / (wall_c + wall_d), "unit": "MB/s", "compress_MBps": mb / wall_c, "decompress_MBps": mb / wall_d,
                    "h2d_bytes_per_step": int(nbytes + r["arc_bytes"]), "d2h_bytes_per_step": int(r["arc_bytes"] + nbytes)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "GB/s of DRAM traffic (ncu dram__bytes per input byte x bytes coded / kernel time)",
                         "traffic_detail": traffic_detail, "peak_source": peak_src,
                         "alg_bytes_per_input_byte": balg,
                         "encode": {"achieved": nbytes * done * balg / code_c / 1e9, "frac": nbytes * done * balg / code_c / 1e9 / peak},
                         "decode": {"achieved": nbytes * done * balg / code_d / 1e9, "frac": nbytes * done * balg / code_d / 1e9 / peak},
                         "kernel": "encode + decode coding kernels (see DESIGN.md section 3)",
                         "kernel_variants": {k: v for k, v in os.environ.items() if k.startswith("ZPQ_") and k != "ZPQ_BENCH_BUDGET_S"},
                         "note": "latency/issue bound, not bandwidth bound (see profiles/ and DESIGN.md section 6)"},
            "transform": None if acc["xf_bytes"] <= 0 else {
                "kernel": "suffix sort of the BWT (zpq_bwt.cu: radix passes + regrouping over the still-unsorted suffixes), rank 0",
                "ms_per_step": acc["xf_s"] * 1e3 / done, "MBps": nbytes * done / 1e6 / acc["xf_s"],
                "alg_bytes_per_input_byte": acc["xf_bytes"] / (nbytes * done),
                "roofline": {"bound": "hbm", "achieved": acc["xf_bytes"] / acc["xf_s"] / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": acc["xf_bytes"] / acc["xf_s"] / 1e9 / peak, "traffic": None}},
            "cpu_baseline": cpu, "bit_exact_vs_reference": bit_exact, "clocks": clocks,
            "archive_ratio": r["arc_bytes"] / nbytes, "setup_s": gen_s, "warmup_s": warm_s, "total_s": time.time() - T_START,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
