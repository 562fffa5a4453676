"""The N>1 host path on CPU: world_size-2 gloo, one flat-bucket gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from odevit_b200.dp import FlatGradAllReduce, shard_batch
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 4)
    extra = torch.nn.Parameter(torch.ones(3))          # never receives a grad on rank 1
    x = torch.arange(48, dtype=torch.float32).reshape(8, 6) / 10
    sl = shard_batch(8, rank, world)
    loss = lin(x[sl]).pow(2).sum() / 8
    if rank == 0:
        loss = loss + extra.sum()
    loss.backward()
    red = FlatGradAllReduce(list(lin.parameters()) + [extra])
    red()
    out[rank] = [lin.weight.grad.clone(), lin.bias.grad.clone(), extra.grad.clone(), red.bucket_bytes]
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 4)
    x = torch.arange(48, dtype=torch.float32).reshape(8, 6) / 10
    (lin(x).pow(2).sum() / 8).backward()
    for r in range(world):
        w, b, e, nbytes = out[r]
        # mean over ranks of the per-shard sums/8 == full-batch gradient / world
        assert torch.allclose(w * world, lin.weight.grad, rtol=1e-5, atol=1e-6)
        assert torch.allclose(b * world, lin.bias.grad, rtol=1e-5, atol=1e-6)
        assert torch.allclose(e, torch.full((3,), 0.5))
        assert nbytes == 4 * (24 + 4 + 3)
    assert torch.equal(out[0][0], out[1][0])


def _worker_sync(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from odevit_b200.dp import FlatGradAllReduce
    torch.manual_seed(100 + rank)                      # replicas that did NOT seed identically
    lin = torch.nn.Linear(5, 3)
    before = lin.weight.detach().clone()
    FlatGradAllReduce(lin.parameters())                # construction broadcasts rank 0's weights
    out[rank] = [before, lin.weight.detach().clone(), lin.bias.detach().clone()]
    dist.destroy_process_group()


def test_construction_broadcasts_rank0_parameters():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_sync, args=(world, _free_port(), out), nprocs=world, join=True)
    assert not torch.equal(out[0][0], out[1][0])       # they started apart
    assert torch.equal(out[0][1], out[0][0])           # rank 0 keeps its own
    assert torch.equal(out[1][1], out[0][1]) and torch.equal(out[1][2], out[0][2])


def test_shard_batch_covers_everything():
    from odevit_b200.dp import shard_batch
    for n in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                sl = shard_batch(n, r, world)
                seen.extend(range(n)[sl])
            assert seen == list(range(n))
