#!/bin/bash
# A/B of one environment switch on the default bench (run under gpurun): ab_generic.sh VAR [tests...]
O=gpurun_out
VAR=$1; shift
if [ $# -gt 0 ]; then
  timeout 900 python -m pytest "$@" -x -q > $O/ab_tests.log 2>&1; echo "tests rc=$?" > $O/ab_rc.txt
  tail -4 $O/ab_tests.log
  grep -q "tests rc=0" $O/ab_rc.txt || exit 1
fi
for v in 0 1 0 1; do
  env $VAR=$v timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 --per-launch ab_pl_$v.txt 2> $O/ab_bench_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$VAR=$v', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline'].get('achieved'),1))"
done
