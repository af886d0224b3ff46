#!/bin/bash
# 8-GPU check of the final build (run under `gpurun --gpus 8`): weak scaling 64 tiles per GPU, then configs[4] (H-G step,
# global batch 256, strong scaling) at 8 GPUs. Each run under its own timeout.
O=gpurun_out
run8() {  # name, args...
  local name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline "$@" > $O/n8_${name}_8gpu.json 2> $O/n8_${name}_8gpu.err; echo "$name rc=$?"
}
run8 train_weak64
run8 hg_strong256 --workload hg --global-batch 256
for f in $O/n8_*_8gpu.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "n", d["n_gpus"], "B/gpu", d["config"]["batch_per_gpu"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 2), d["scaling"], d["clocks"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
