#!/bin/bash
# r02t: after the decompress fix (PCOMP blocks staged at once) and the dec7 instantiation for the BWT model
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02t_session.log
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,serial --format=csv,noheader
echo "== pytest transforms + archive (production tier) + smoke"
( time timeout 420 python -m pytest tests/test_gpu_transforms.py tests/test_gpu_archive.py -m gpu -x -q -k "not dec6 and not dec8 and not fast and not generic" 2>&1 | tail -4 ) 2>&1
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench --config bwt_256"
timeout 300 python bench.py --config bwt_256 --steps 3 --warmup 2 > $OUT/r02t_bench_bwt_256.json 2> $OUT/r02t_bench_err.log; cat $OUT/r02t_bench_bwt_256.json; tail -3 $OUT/r02t_bench_err.log
echo "== done"
} > $L 2>&1
tail -40 $L
