#!/bin/bash
# DRAM traffic and tensor-pipe activity of every tcgen05 launch of ONE adversarial train step (run under gpurun,
# after the same bench command has exited 0 without ncu). The first three (warm-up) steps are skipped.
# usage: tools/ncu_tensorcore_step.sh <batch> <tag> <tensor-core launches per step>
set -e
B=${1:-64}; TAG=${2:-r01}; N=${3:-102}
CMD="python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none -k 'regex:conv_igemm_kernel|conv_halo|wgrad_igemm|wgrad_halo|rowgemm64|tapdot|tapwgrad_kernel' \
    --launch-skip $((3 * N)) --launch-count $N --csv --log-file gpurun_out/tc_metrics_$TAG.csv $CMD > gpurun_out/ncu_tc_$TAG.log 2>&1
