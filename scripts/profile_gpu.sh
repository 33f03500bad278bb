#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): per-launch device times of the bench run and
# one full-set capture of the dominant kernel.  Run under gpurun; outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ARGS2="--steps 1 --warmup 0 --no-cpu-baseline"
python bench.py $ARGS2 > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_full -c 2 -o gpurun_out/prof_full \
    python bench.py $ARGS2 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
