#!/bin/bash
# r02u: the shim after BWT / E8E9 methods were routed to the GPU library: its GPU tests (zpaq add -method 3/4/5, zpaq x, stream API)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02u_session.log
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,serial --format=csv,noheader
echo "== pytest tests/test_shim.py -m gpu"
( time timeout 400 python -m pytest tests/test_shim.py -m gpu -x -q 2>&1 | tail -6 ) 2>&1
echo "== zpaq add -method 3 on 64 MiB of text through the shim (BWT on the device), -threads 64, next to the stock CLI"
W=$(mktemp -d /tmp/zpq3.XXXXXX); mkdir $W/data
python - "$W" <<'PY'
import sys, os
sys.path.insert(0, "oracle")
import oraclepy as O
w = sys.argv[1]
with open(os.path.join(w, "data", "t.txt"), "wb") as fh:
    for k in range(64):
        fh.write(O.gen_block(0, 8000 + k, 1 << 20))
PY
cd $W
( time /root/repo/oracle/_ref/zpaq_ref add r.zpaq data -method 30 -threads 16 2>&1 | tail -1 ) 2>&1 | grep -E "real|OK"
( time ZPQ_SHIM_LOG=1 /root/repo/oracle/_ref/zpaq_b200 add g.zpaq data -method 30 -threads 64 2> shim.log | tail -1 ) 2>&1 | grep -E "real|OK"
grep -c -- "-> GPU, status 0" shim.log; grep -c "reference CPU code" shim.log
ls -l r.zpaq g.zpaq | awk '{print $5, $9}'
/root/repo/oracle/_ref/zpaq_ref x g.zpaq -to og > /dev/null 2>&1; cmp data/t.txt og/data/t.txt && echo "stock zpaq extracts the GPU-written BWT archive: identical"
cd /root/repo; rm -rf $W
echo "== done"
} > $L 2>&1
tail -40 $L
