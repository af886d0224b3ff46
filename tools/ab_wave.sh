#!/bin/bash
# A/B of the whole-wave grids and the 3-CTA BatchNorm-backward reduce (run under gpurun): isolated kernel timings,
# GPU tests, alternating bench lines, ncu launch list of the default build.
O=gpurun_out
for cfg in "" "TG_WAVE_GRID=0" "TG_BN_REDUCE_OCC=1" ; do
  echo "== time_bw [$cfg]" >> $O/ab_wave_bw.txt
  env $cfg python tools/time_bw.py 64 >> $O/ab_wave_bw.txt 2>&1
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/ab_wave_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/ab_wave_tests.log
for i in 1 2; do
  for cfg in "TG_WAVE_GRID=0 TG_BN_REDUCE_OCC=1" "TG_WAVE_GRID=1" ; do
    env $cfg python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>> $O/ab_wave_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> $O/ab_wave_bench.txt
  done
done
cat $O/ab_wave_bench.txt
bash tools/launchlist.sh > /dev/null 2>&1
cat $O/ab_wave_bw.txt
