#!/bin/bash
# tools/r02_session10.sh -- lock-step decoder after the register / load-hoisting pass on the chain, mix and tail roles.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02j}
L=$OUT/${R}_session.log
echo "== parity" | tee $L
timeout 300 python tools/dec_check.py 2>&1 | tail -2 | tee -a $L
echo "== timing 1029 x 32 KiB, method 5: dec8 / dec7" | tee -a $L
for k in 8 7; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 2>&1 | grep "rep 1" | tee -a $L; done
echo "== method 4" | tee -a $L
for k in 8 7; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 4 2>&1 | grep "rep 1" | tee -a $L; done
echo "== ncu dec8" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_kernel8 -c 1 -o $OUT/${R}_dec8 -f python tools/prof_case.py 5 8 16384 both > $OUT/${R}_ncu_dec.log 2>&1
tail -2 $OUT/${R}_ncu_dec.log | tee -a $L
echo "== done" | tee -a $L
