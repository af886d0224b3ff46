#!/bin/bash
# 2-GPU check (run under `gpurun --gpus 2`): NCCL gradient-parity test, then weak scaling 1 -> 2 GPUs with the default bench.
O=gpurun_out
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py -x -q -m gpu > $O/n2_ddp_test.log 2>&1; echo "ddp test rc=$?"; tail -3 $O/n2_ddp_test.log
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > $O/n2_weak_1gpu.json 2> $O/n2_weak_1gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
  bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $O/n2_weak_2gpu.json 2> $O/n2_weak_2gpu.err
for f in $O/n2_weak_1gpu.json $O/n2_weak_2gpu.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "n", d["n_gpus"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 2), d["clocks"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
