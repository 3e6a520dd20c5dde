"""World-size-2 gloo test (CPU) of the event-sharded data-parallel plumbing in sparseeventid_b200/trainer.py:
the flat gradient arena + bucketed all-reduce must give every rank the MEAN of the per-rank gradients, i.e.
exactly what the reference's DDP / Horovod allreduce(average) produces."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                               torch.nn.Linear(16, 3))


def _worker(rank, world, port, q, overlap=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparseeventid_b200.trainer import FlatGradArena
    model = _model()
    if rank == 1:                       # rank 0's weights must win the initial broadcast
        for p in model.parameters():
            p.data.add_(1.0)
    for p in model.parameters():
        dist.broadcast(p.data, src=0)
    arena = FlatGradArena(list(model.parameters()), bucket_bytes=256, overlap=overlap)   # tiny buckets -> several all-reduces
    assert len(arena.buckets) > 2
    g = torch.Generator().manual_seed(100 + rank)
    for step in range(2):
        x = torch.randn(5, 6, generator=g)
        arena.zero()
        model(x).pow(2).sum().backward()
        arena.finish()
    # numpy (pickled by value): tensors would travel by file descriptor and need this process alive at receive time
    q.put((rank, [p.grad.clone().numpy() for p in model.parameters()], [p.data.clone().numpy() for p in model.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_flat_arena_allreduce_mean_world2(overlap):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, overlap)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, grads, params = q.get(timeout=120)
        res[r] = ([torch.from_numpy(g) for g in grads], [torch.from_numpy(v) for v in params])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: each rank's own gradient on its own shard, averaged
    per_rank = []
    for rank in range(world):
        model = _model()
        g = torch.Generator().manual_seed(100 + rank)
        for step in range(2):
            x = torch.randn(5, 6, generator=g)
            model.zero_grad()
            model(x).pow(2).sum().backward()
        per_rank.append([p.grad.clone() for p in model.parameters()])
    mean = [sum(gs) / world for gs in zip(*per_rank)]
    for rank in range(world):
        for got, want in zip(res[rank][0], mean):
            assert torch.allclose(got, want, atol=1e-6)
        for a, b in zip(res[rank][1], res[0][1]):
            assert torch.equal(a, b)            # same (rank-0) parameters everywhere


def test_arena_single_process_is_plain_backward():
    from sparseeventid_b200.trainer import FlatGradArena
    model = _model()
    arena = FlatGradArena(list(model.parameters()))
    x = torch.randn(4, 6)
    arena.zero()
    model(x).sum().backward()
    arena.finish()
    ref = _model()
    ref(x).sum().backward()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a.grad, b.grad)
        assert a.grad.data_ptr() >= arena.flat.data_ptr()      # gradients live in the flat arena


def test_arena_rebinds_gradients_after_zero_grad_set_to_none():
    """optimizer.zero_grad() defaults to set_to_none=True: the next arena.zero() must point .grad back into the arena."""
    import torch
    from sparseeventid_b200.trainer import FlatGradArena
    lin = torch.nn.Linear(4, 3)
    arena = FlatGradArena(list(lin.parameters()))
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    arena.zero()
    lin(torch.ones(2, 4)).sum().backward()
    assert float(arena.flat.abs().sum()) > 0
    opt.zero_grad()                                         # set_to_none=True
    assert lin.weight.grad is None
    arena.zero()
    assert lin.weight.grad is not None and lin.weight.grad.data_ptr() >= arena.flat.data_ptr()
    lin(torch.ones(2, 4)).sum().backward()
    assert float(arena.flat.abs().sum()) > 0
    lo, hi = arena.flat.data_ptr(), arena.flat.data_ptr() + 4 * arena.flat.numel()
    assert all(lo <= p.grad.data_ptr() < hi for p in lin.parameters())
