#!/bin/bash
# Multi-GPU pass: bash scripts/gpu_multi.sh <tag> <ngpus> [local qubits]
TAG=${1:-m}; N=${2:-2}; LQ=${3:-28}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpus_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > $OUT/pytest_dist_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_dist_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
  bench.py --gpus $N --steps 2 --warmup 3 --sharded-local-qubits $LQ > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench rc=$?"; cat $OUT/bench_${TAG}_n$N.json; tail -5 $OUT/bench_${TAG}_n$N.err
