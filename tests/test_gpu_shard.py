"""GPU: the sharded data path over NCCL -- raw blocks scattered from rank 0 into the ranks' HBM, coded by
zpq_compress_blocks_dev where they land, coded blocks gathered; then the archive scattered, decoded by
zpq_decode_blocks_dev, verified on the device and gathered.  Runs with as many ranks as the box has GPUs (at most 2);
on a one-GPU box the same code runs with world size 1."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_sharded_roundtrip_over_nccl_matches_single_gpu_archive():
    import torch
    n = min(2, torch.cuda.device_count())
    assert n >= 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "sharded_roundtrip.py"), "--blocks", "24", "--bytes", "150000", "--check-single"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["roundtrip_ok"] and res["archive_equals_single_gpu"] and res["n_gpus"] == n
