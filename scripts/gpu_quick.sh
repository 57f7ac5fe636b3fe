#!/bin/bash
# Quick GPU pass: parity tests + bench (no profiler).  bash scripts/gpu_quick.sh <tag> [bench args]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
python bench.py "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
