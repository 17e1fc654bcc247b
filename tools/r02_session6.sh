#!/bin/bash
# tools/r02_session6.sh -- wave encoder with one shared (non-inlined) recipe interpreter: variants, group sizes, ncu.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02f}
L=$OUT/${R}_session.log
echo "== encoder variants, 1029 x 64 KiB, chunk 512 (shared interpreter build)" | tee $L
for v in 41 9 57 25 33; do echo "-- ZPQ_WAVE_VARIANT=$v" | tee -a $L; ZPQ_WAVE_VARIANT=$v timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
for c in 256 1024; do echo "-- variant 41 chunk $c" | tee -a $L; ZPQ_WAVE_VARIANT=41 ZPQ_CHUNK=$c timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
for g in 8 32; do echo "-- variant 41 group $g" | tee -a $L; ZPQ_WAVE_VARIANT=41 ZPQ_GROUP=$g timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
echo "-- variant 41 per-role cycles" | tee -a $L
ZPQ_WAVE_VARIANT=41 ZPQ_WAVE_PROF=1 timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L
echo "== ncu wave variant 41" | tee -a $L
ZPQ_WAVE_VARIANT=41 timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_wave_kernel -s 60 -c 1 -o $OUT/${R}_wave -f python tools/prof_case.py 5 1029 65536 enc > $OUT/${R}_ncu_wave.log 2>&1
tail -2 $OUT/${R}_ncu_wave.log | tee -a $L
echo "== done" | tee -a $L
