#!/bin/bash
# tools/r02_session2.sh -- first run of the critical-first decoder (zpq_kernels7.cu) on a B200: parity on a small batch,
# timing against the round-1 decoder, ncu --set full at the benchmark's per-SM occupancy.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02b}
L=$OUT/${R}_session.log
echo "== quick parity (kernel7 vs oracle), methods 4 / 5 / chain17" | tee $L
timeout 300 python tools/dec_check.py 2>&1 | tail -20 | tee -a $L
echo "== decoder timing 1029 x 32 KiB: dec7 vs dec6" | tee -a $L
for k in 7 6; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 2>&1 | tee -a $L; done
echo "== method 4, 1029 x 32 KiB" | tee -a $L
for k in 7 6; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 4 2>&1 | tee -a $L; done
echo "== ncu dec7" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_kernel7 -c 1 -o $OUT/${R}_dec7 -f python tools/prof_case.py 5 7 16384 both > $OUT/${R}_ncu_dec.log 2>&1
tail -3 $OUT/${R}_ncu_dec.log | tee -a $L
echo "== gpu tests (archive, parity)" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_archive.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
echo "== done" | tee -a $L
