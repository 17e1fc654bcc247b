#!/bin/bash
# r02r: pre-processing kernels on hardware (parity, timing), the CLI through the shim with the arrival-following window
cd "$(dirname "$0")/.."
L=gpurun_out/r02r_session.log
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,serial --format=csv,noheader
echo "== pytest transforms + shim"
timeout 900 python -m pytest tests/test_gpu_transforms.py tests/test_shim.py -x -q -m gpu 2>&1 | tail -8
echo "== BWT timing"
timeout 600 python tools/bwt_timing.py 256 2>&1 | tail -12
echo "== CLI, window 100 ms"
TS=512 SKIP_SUBSET=1 timeout 900 bash tools/cli_threads.sh 512 2>&1
echo "== done"
} > $L 2>&1
tail -60 $L
