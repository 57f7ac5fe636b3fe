#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list + full capture of the sweep kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; free -g >> $OUT/gpu_$TAG.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
BCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-mps"
$BCMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv $BCMD > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
$BCMD > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sv_sweep -s 60 -c 3 -f -o $OUT/prof_sweep_$TAG $BCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
