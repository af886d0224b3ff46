"""Multi-GPU numerical parity: after the bucketed NCCL allreduce every .grad equals the mean over
ranks of the per-shard oracle gradients, and the replicas stay bit-identical after the optimizer steps. Needs >= 2
GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`, log kept under profiles/)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["tf32x3", "bf16"])
def test_two_rank_nccl_gradient_parity(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "ddp_parity.py"), "--mode", mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0 and lines, res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"] and out["world"] == 2 and out["tensors"] == 58 + 16
