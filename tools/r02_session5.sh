#!/bin/bash
# tools/r02_session5.sh -- decoder after the instruction diet; encoder variants vs instruction-cache footprint.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02e}
L=$OUT/${R}_session.log
echo "== decoder parity + timing (dec7 after the diet vs dec6), 1029 x 32 KiB" | tee $L
timeout 300 python tools/dec_check.py 2>&1 | tail -3 | tee -a $L
for k in 7 6; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 2>&1 | grep "rep 1" | tee -a $L; done
echo "== encoder variants at chunk 512 (hot code size: 25 = unrolled ICM/ISSE/SSE loops, 9 = one generic loop)" | tee -a $L
for v in 25 57 9 41; do echo "-- ZPQ_WAVE_VARIANT=$v" | tee -a $L; ZPQ_WAVE_VARIANT=$v timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
echo "-- variant 9 per-role cycles" | tee -a $L
ZPQ_WAVE_VARIANT=9 ZPQ_WAVE_PROF=1 timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L
echo "== wide models, prep kernels" | tee -a $L
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_prep.py -x -q -m gpu 2>&1 | tail -3 | tee -a $L
echo "== ncu dec7 (after the diet)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_kernel7 -c 1 -o $OUT/${R}_dec7 -f python tools/prof_case.py 5 7 16384 both > $OUT/${R}_ncu_dec.log 2>&1
tail -2 $OUT/${R}_ncu_dec.log | tee -a $L
echo "== done" | tee -a $L
