#!/bin/bash
# Quick state of the current build (run under gpurun): GPU tests, default bench line with the per-launch table, ncu launch list.
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/qs_tests.log 2>&1; echo "tests rc=$?" > $O/qs_rc.txt
python bench.py --no-cpu-baseline --steps 8 --warmup 3 --per-launch qs_per_launch.txt > $O/qs_bench.json 2> $O/qs_bench.err; echo "bench rc=$?" >> $O/qs_rc.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/qs_launches.csv python tools/profile_step.py 64 3 > $O/qs_ncu.log 2>&1; echo "ncu rc=$?" >> $O/qs_rc.txt
python tools/summarize_launches.py $O/qs_launches.csv > $O/qs_launches_summary.txt 2>&1
tail -3 $O/qs_tests.log; cat $O/qs_rc.txt; cat $O/qs_bench.json | head -c 600
