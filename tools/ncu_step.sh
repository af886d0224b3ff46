#!/bin/bash
# Launch list (device time of every kernel) of one adversarial train step — run under gpurun.
# usage: tools/ncu_step.sh <batch> <tag>
set -e
B=${1:-16}; TAG=${2:-r01}
CMD="python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
