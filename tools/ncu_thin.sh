#!/bin/bash
# `ncu --set full` captures of the thin (1 <-> 64 channel) kernels of one B=64 adversarial step
set -e
python tools/profile_step.py 64 3 > gpurun_out/thin_plain.log 2>&1
for k in 'rowgemm64_kernel' 'tapdot_kernel' 'tapsum_kernel' 'tapwgrad_kernel'; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$k" -c 4 \
      -o gpurun_out/thin_full_$k -f python tools/profile_step.py 64 3 > gpurun_out/thin_ncu_$k.log 2>&1 || true
done
ls -la gpurun_out/*.ncu-rep
