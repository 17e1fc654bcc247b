#!/bin/bash
# tools/r02_session3.sh -- encoder warp order A/B, then the reworked bench.py end to end (budget guard, reference arm).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02c}
L=$OUT/${R}_session.log
echo "== encoder 1029 x 64 KiB: role-major (default) vs group-major, chunk 512/2048" | tee $L
for o in 1 0; do for c in 2048 512; do echo "-- ZPQ_WAVE_ORDER=$o ZPQ_CHUNK=$c" | tee -a $L; ZPQ_WAVE_ORDER=$o ZPQ_CHUNK=$c timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done; done
echo "-- role-major, group 32, chunk 512" | tee -a $L
ZPQ_GROUP=32 ZPQ_CHUNK=512 timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L
echo "-- role-major per-role cycles" | tee -a $L
ZPQ_WAVE_PROF=1 ZPQ_CHUNK=512 timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L
echo "== reference arm (3 steps)" | tee -a $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${R}_bench_ref.json 2>> $L; tail -c 1200 $OUT/${R}_bench_ref.json | tee -a $L
echo "== bench (3 steps, 3 warm-up)" | tee -a $L
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/${R}_bench_1gpu.json 2>> $L; tail -c 3000 $OUT/${R}_bench_1gpu.json | tee -a $L
echo "== done" | tee -a $L
