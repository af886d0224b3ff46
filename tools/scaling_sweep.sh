#!/bin/bash
# Multi-GPU sweep on ONE 8-GPU box (run under `gpurun --gpus 8`): BASELINE.json configs[3] (train step, global batch 512
# over 2/4/8 GPUs), configs[4] (human-guided step, global batch 256, strong scaling over 1/2/4/8, M_large masks) and
# the weak-scaling headline at 8 GPUs. One JSON line per run -> gpurun_out/<tag>_*.json
TAG=${1:-r02}
run() {  # n, name, args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_${n}gpu.json 2> gpurun_out/${TAG}_${name}_${n}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_${n}gpu.json 2> gpurun_out/${TAG}_${name}_${n}gpu.err
  fi
  tail -c 300 gpurun_out/${TAG}_${name}_${n}gpu.json | head -c 10 > /dev/null
}
for n in 1 2 4 8; do run $n hg_strong256 --workload hg --global-batch 256; done
for n in 2 4 8; do run $n train_global512 --global-batch 512; done
run 8 train_weak64
run 1 train_weak64
for f in gpurun_out/${TAG}_*gpu.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "n", d["n_gpus"], "B/gpu", d["config"]["batch_per_gpu"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 2), d["scaling"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
