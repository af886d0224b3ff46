#!/bin/bash
# Streaming-kernel timings + GPU tests + ncu launch list of the current build (run under gpurun).
O=gpurun_out
python tools/time_bw.py 64 > $O/quick_bw.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/quick_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/quick_tests.log
bash tools/launchlist.sh > /dev/null 2>&1
python bench.py --no-cpu-baseline --steps 10 --warmup 3 2> $O/quick_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1), 'pconv', round(d['roofline']['pconv_only']['tflops'],1), 'frac', round(d['roofline']['frac'],3))"
cat $O/quick_bw.txt
grep "upsample\|total" $O/ll_launches_summary.txt
