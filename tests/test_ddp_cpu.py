"""CPU, world_size 2 over gloo: the data-parallel plumbing (episode sharding + flat-bucket gradient all-reduce)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optimalstrategiesagainstgenerativeattacks_b200 import ddp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2))]
        g = torch.Generator().manual_seed(100 + rank)
        params[0].grad = torch.randn(5, 3, generator=g)
        params[1].grad = torch.randn(7, generator=g)      # params[2] never gets a gradient -> must be left out
        bucket = ddp.FlatGradBucket(params)
        assert bucket.flat.numel() == 22 and params[2].grad is None
        assert params[0].grad.data_ptr() == bucket.flat.data_ptr()
        bucket.all_reduce()
        expect0 = sum(torch.randn(5, 3, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
        assert torch.allclose(params[0].grad, expect0)
        lo, hi = ddp.shard_range(8, rank, world)
        out[rank] = (lo, hi, float(bucket.flat.sum()))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29731, out), nprocs=world, join=True)
    assert out[0][:2] == (0, 4) and out[1][:2] == (4, 8)
    assert abs(out[0][2] - out[1][2]) < 1e-6


def test_shard_range_rejects_ragged_batches():
    with pytest.raises(ValueError):
        ddp.shard_range(10, 0, 4)
    assert [ddp.shard_range(128, r, 8) for r in (0, 7)] == [(0, 16), (112, 128)]


def _worker_attach(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                       # replicas start DIFFERENT: attach() must broadcast rank 0's parameters
        net = torch.nn.Linear(4, 3)
        ddp.broadcast_module_state(net)
        opt = torch.optim.SGD(net.parameters(), lr=0.5)
        ddp.attach(opt, defer=True)                   # CPU tensors: the deferred mode degrades to reduce-then-step, flush() is a no-op
        w0 = net.weight.detach().clone()
        x = torch.full((2, 4), float(rank + 1))
        net(x).sum().backward()
        g_local = net.weight.grad.detach().clone()
        opt.step()
        assert opt.flush() is None
        gathered = [torch.zeros_like(g_local) for _ in range(world)]
        dist.all_gather(gathered, g_local)
        mean_g = sum(gathered) / world
        assert torch.allclose(net.weight, w0 - 0.5 * mean_g, atol=1e-6)       # the step used the mean over the global batch
        ws = [torch.zeros_like(net.weight) for _ in range(world)]
        dist.all_gather(ws, net.weight.detach())
        assert torch.equal(ws[0], ws[1])
        # a parameter that starts receiving gradients later must not be skipped silently
        extra = torch.nn.Parameter(torch.ones(2))
        opt.add_param_group({"params": [extra]})
        extra.grad = torch.ones(2)
        try:
            opt.step()
            out[rank] = "no error"
        except RuntimeError as e:
            out[rank] = "raised" if "changed" in str(e) else str(e)
        del mean_g
    finally:
        dist.destroy_process_group()


def test_attach_broadcasts_reduces_and_detects_new_gradients():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_attach, args=(world, 29741, out), nprocs=world, join=True)
    assert out[0] == "raised" and out[1] == "raised"
