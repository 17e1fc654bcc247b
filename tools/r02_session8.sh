#!/bin/bash
# tools/r02_session8.sh -- encoder variants at the benchmark's size (1029 x 1 044 480 B), two repetitions each, interleaved order;
# full-size mixed-model parity test.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02h}
L=$OUT/${R}_session.log
echo "== encoder variants, 1029 x 1044480 B (benchmark size)" | tee $L
timeout 900 python tools/enc_variants.py 1029 1044480 41,25,57,41,25,57 2>&1 | tee -a $L
echo "== full-size mixed-model parity test" | tee -a $L
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size" 2>&1 | tail -15 | tee -a $L
echo "== done" | tee -a $L
