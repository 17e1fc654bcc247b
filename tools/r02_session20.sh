#!/bin/bash
# r02v: store / LZ77 paths on hardware: parity tests of all pre-processing paths, LZ77 throughput
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02v_session.log
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,serial --format=csv,noheader
echo "== pytest tests/test_gpu_transforms.py"
( time timeout 200 python -m pytest tests/test_gpu_transforms.py -m gpu -x -q 2>&1 | tail -5 ) 2>&1
echo "== LZ77 / store timing, 128 x 1 MiB"
timeout 60 python tools/lz_timing.py 128 2>&1 | tail -5
echo "== done"
} > $L 2>&1
tail -30 $L
