#!/bin/bash
# Launch list (device time of every kernel) of ONE adversarial train step — run under gpurun.
# The first three (warm-up) steps are skipped with --launch-skip so that only one step is replayed under ncu.
# usage: tools/ncu_step.sh <batch> <tag> [launches_per_step]
set -e
B=${1:-16}; TAG=${2:-r01}; LPS=${3:-540}
CMD="python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3 * LPS)) --launch-count $LPS --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
