#!/bin/bash
# tools/r02_session4.sh -- 2 GPUs: device-resident analysis + sharded NCCL data path tests, bench.py under torchrun.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02d}
L=$OUT/${R}_session.log
nvidia-smi -L | tee $L
echo "== prep + shard tests" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_prep.py tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -8 | tee -a $L
echo "== sharded round trip, 2 ranks, 128 full-size blocks" | tee -a $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_roundtrip.py --blocks 128 --check-single 2>&1 | grep -v Warning | tail -3 | tee -a $L
echo "== bench --mode sharded --gpus 2 (256 blocks per GPU)" | tee -a $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode sharded --blocks 256 --steps 1 --warmup 1 > $OUT/${R}_bench_sharded_2gpu.json 2>> $L; tail -c 2500 $OUT/${R}_bench_sharded_2gpu.json | tee -a $L
echo "== bench --gpus 2 (contract line), 2 steps" | tee -a $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/${R}_bench_2gpu.json 2>> $L; tail -c 2500 $OUT/${R}_bench_2gpu.json | tee -a $L
echo "== done" | tee -a $L
