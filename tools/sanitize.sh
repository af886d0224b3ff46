#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (run under gpurun; logs -> gpurun_out/, summaries kept in profiles/).
# memcheck: out-of-bounds / misaligned accesses of every kernel family; synccheck: barrier misuse; racecheck: shared-memory
# hazards of the CUDA-core kernels (the tcgen05 / TMA kernels synchronise through mbarriers and the async proxy, which
# racecheck does not model: they are covered by memcheck + synccheck + the numerical tests).
TAG=${1:-r02}
SEL='tests/test_elementwise_gpu.py tests/test_conv_igemm_gpu.py'
export PYTHONDONTWRITEBYTECODE=1
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 0 --log-file gpurun_out/${TAG}_sanitizer_${tool}.log \
     python -m pytest $SEL tests/test_aux_gpu.py -x -q -k "not full_size and not batched_inpainter and not evaluate_drop_in" > gpurun_out/${TAG}_sanitizer_${tool}.pytest.log 2>&1
  echo "$tool exit $?" >> gpurun_out/${TAG}_sanitizer_summary.txt
  tail -3 gpurun_out/${TAG}_sanitizer_${tool}.log >> gpurun_out/${TAG}_sanitizer_summary.txt
  tail -2 gpurun_out/${TAG}_sanitizer_${tool}.pytest.log >> gpurun_out/${TAG}_sanitizer_summary.txt
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 --log-file gpurun_out/${TAG}_sanitizer_racecheck.log \
   python -m pytest tests/test_elementwise_gpu.py tests/test_aux_gpu.py -x -q -k "not full_size and not batched_inpainter and not evaluate_drop_in" > gpurun_out/${TAG}_sanitizer_racecheck.pytest.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_sanitizer_summary.txt
tail -3 gpurun_out/${TAG}_sanitizer_racecheck.log >> gpurun_out/${TAG}_sanitizer_summary.txt
tail -2 gpurun_out/${TAG}_sanitizer_racecheck.pytest.log >> gpurun_out/${TAG}_sanitizer_summary.txt
cat gpurun_out/${TAG}_sanitizer_summary.txt
