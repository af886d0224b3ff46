#!/bin/bash
# ncu launch list of one B=64 adversarial step (run under gpurun) -> gpurun_out/ll_launches_summary.txt
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/ll_launches.csv python tools/profile_step.py 64 3 > $O/ll_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $O/ll_launches.csv > $O/ll_launches_summary.txt 2>&1
head -30 $O/ll_launches_summary.txt
