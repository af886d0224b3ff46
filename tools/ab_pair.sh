#!/bin/bash
# A/B of the CTA-pair conv kernel (run under gpurun): parity tests first, then the default bench with and without it.
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_pair_gpu.py -x -q > $O/pair_tests.log 2>&1; echo "tests rc=$?" > $O/pair_rc.txt
tail -5 $O/pair_tests.log
if grep -q "tests rc=0" $O/pair_rc.txt; then
  for v in 0 1 0 1; do
    TG_NO_CONV_PAIR=$v timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 --per-launch pair_pl_$v.txt 2> $O/pair_bench_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nopair=$v', d['value'], d['ms_per_step'], d['clocks'], d['roofline'].get('achieved'))"
  done
fi
