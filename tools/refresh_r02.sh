#!/bin/bash
# Re-measure the 1-GPU round-2 records with the current build (run under gpurun): bench lines, per-launch table,
# ncu launch list of one step, tensor-core metrics of one step.
set -e
O=gpurun_out
python bench.py --impl reference > $O/r02_bench_reference_arm.json 2> $O/ref_arm.err
python bench.py > $O/r02_bench_b64_1gpu.json 2> $O/bench.err
python bench.py --workload infer --batch 16 --no-cpu-baseline > $O/r02_bench_infer_b16_1gpu.json 2>> $O/bench.err
python bench.py --workload infer --batch 16 --no-graph --no-cpu-baseline > $O/r02_bench_infer_b16_1gpu_nograph.json 2>> $O/bench.err
python bench.py --workload hg --no-cpu-baseline > $O/r02_bench_hg_b64_1gpu.json 2>> $O/bench.err
python bench.py --no-cpu-baseline --steps 8 --warmup 3 --per-launch r02_tensorcore_per_launch_b64.txt > $O/perlaunch_line.json 2>> $O/bench.err
python tools/profile_step.py 64 3 > $O/r02_profile_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/r02_launches_b64_step.csv python tools/profile_step.py 64 3 > $O/r02_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none --profile-from-start off \
    -k 'regex:conv_igemm_kernel|conv_halo|wgrad_igemm|wgrad_wide|wgrad_halo|rowgemm64|tapdot|tapwgrad_kernel' \
    --csv --log-file $O/r02_ncu_tensorcore_step_metrics.csv python tools/profile_step.py 64 3 > $O/r02_ncu_tc.log 2>&1
# --set full captures: the first three conv_igemm launches of the step are enc2 (<128, kMT = 2>), enc3 and enc4 (the CTA-pair <256> kernel)
for k in 'conv_igemm_kernel' 'conv_halo_kernel' 'wgrad_wide_kernel' 'bn_bwd_reduce_kernel'; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$k" -c 3 \
      -o $O/r02c_full_$k -f python tools/profile_step.py 64 3 > $O/r02c_ncu_full_$k.log 2>&1 || true
done
python tools/time_bw.py 64 > $O/r02_streaming_kernels_timing.txt 2>&1
ls -la $O/*.ncu-rep
