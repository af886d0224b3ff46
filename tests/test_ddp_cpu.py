"""The data-parallel gradient exchange (tg_b200.ddp) on CPU: gloo backend, world_size 2.
Checks bucketing, the flush order, averaging into .grad and multi-contribution parameters."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, bucket_bytes, q):
    sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tg_b200 import functional
    from tg_b200.ddp import BucketedGradReducer, broadcast_module_state

    torch.manual_seed(rank)                      # different initial weights per rank -> broadcast must fix
    m = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    broadcast_module_state([m])
    w0 = m[0].weight.detach().clone()
    red = BucketedGradReducer([m], bucket_bytes=bucket_bytes)
    names = [n for n, _ in m.named_parameters()]
    # emulate what the engines do inside backward: per-"layer" gradients announced in reverse order,
    # the last layer twice (a parameter with two contributions in one backward, like D in the D step)
    grads = {n: torch.full_like(p, float(rank + 1)) * (i + 1) for i, (n, p) in enumerate(m.named_parameters())}
    extra = torch.full_like(m[1].weight, 10.0 * (rank + 1))
    hook = functional._hook_for(m)
    hook(["1.weight", "1.bias"], [grads["1.weight"], grads["1.bias"]])
    hook(["1.weight"], [extra])
    hook(["0.weight", "0.bias"], [grads["0.weight"], grads["0.bias"]])
    for n, p in m.named_parameters():            # what autograd would have accumulated locally
        p.grad = grads[n].clone() + (extra if n == "1.weight" else 0)
    red.finish()
    avg = (1 + world) / 2.0
    ok = True
    for i, (n, p) in enumerate(m.named_parameters()):
        expect = avg * (i + 1) + (10.0 * avg if n == "1.weight" else 0.0)
        ok = ok and torch.allclose(p.grad, torch.full_like(p, expect))
    q.put((rank, ok, red.buckets_launched, w0))
    red.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [1, 1 << 20])
def test_bucketed_reducer_gloo_world2(bucket_bytes):
    world, port = 2, 29500 + os.getpid() % 500 + (0 if bucket_bytes == 1 else 501)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bucket_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "averaged gradients differ from the expected rank mean"
    assert torch.equal(res[0][3], res[1][3]), "initial broadcast did not equalise the replicas"
    assert res[0][2] == (3 if bucket_bytes == 1 else 1)      # tiny buckets: one per announcement; big: one flush
