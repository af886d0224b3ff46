#!/bin/bash
# Round-2 ncu captures of ONE adversarial train step at B=64 (tools/profile_step.py brackets the step with
# cudaProfilerStart/Stop; the same program has exited 0 without ncu first). Run under gpurun, one GPU.
#   1. metrics pass over every tensor-core launch of the step: duration, DRAM bytes, tensor-pipe activity
#   2. `--set full` captures (with source) of the top kernels -> .ncu-rep files read back with `ncu -i`
set -e
python tools/profile_step.py 64 3 > gpurun_out/r02_profile_plain.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none --profile-from-start off \
    -k 'regex:conv_igemm_kernel|conv_halo|wgrad_igemm|wgrad_halo|rowgemm64|tapdot|tapwgrad_kernel' \
    --csv --log-file gpurun_out/r02_ncu_tensorcore_step_metrics.csv python tools/profile_step.py 64 3 > gpurun_out/r02_ncu_tc.log 2>&1
for k in 'conv_halo_kernel' 'conv_igemm_kernel' 'wgrad_igemm_kernel' 'bn_bwd_reduce_kernel' 'conv_halo_stream_kernel'; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$k" -c 2 \
      -o gpurun_out/r02_full_$k -f python tools/profile_step.py 64 3 > gpurun_out/r02_ncu_full_$k.log 2>&1 || true
done
ls -la gpurun_out/*.ncu-rep
