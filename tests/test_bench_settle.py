"""bench.py's allocator-settling loop (setup before the timed loops): every step contains the gradient all-reduce, so
the ranks must run the same number of steps even when their allocators go quiet at different times.  World-size-2 gloo
run on CPU: the ranks see allocations at different steps, each step all-reduces; a rank-local stopping rule would leave
one rank waiting in a collective the other never enters (the N=2 hang of the first version of this loop)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_settle_single_rank_counts():
    import bench
    allocs = {"n": 0}
    noisy = {0, 1, 4}                    # steps that allocate

    def step(i):
        if i in noisy:
            allocs["n"] += 1

    n = bench.settle_allocator(step, lambda: allocs["n"], lambda f: f, chunk=3)
    assert n == 9                        # chunks [0-2], [3-5] allocate, [6-8] is the first quiet one
    allocs["n"] = 0
    n = bench.settle_allocator(lambda i: allocs.__setitem__("n", allocs["n"] + 1), lambda: allocs["n"], lambda f: f,
                               chunk=3, cap=11)
    assert n == 12                       # never quiet: the cap ends it (after the chunk that crosses it)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    allocs = {"n": 0}
    noisy = {0: {0, 1}, 1: {0, 2, 4}}[rank]          # rank 1's allocator settles later (alone, rank 0 would stop after 6)
    seen = []

    def step(i):
        if i in noisy:
            allocs["n"] += 1
        t = torch.tensor([float(i + rank)])           # the step's collective
        dist.all_reduce(t)
        seen.append(float(t.item()))

    def any_rank(flag):
        t = torch.tensor([1.0 if flag else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(t.item() > 0)

    n = bench.settle_allocator(step, lambda: allocs["n"], any_rank, chunk=3)
    q.put((rank, n, seen))
    dist.barrier()
    dist.destroy_process_group()


def test_settle_ranks_agree_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, n, seen = q.get(timeout=120)
        res[r] = (n, seen)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == res[1][0] == 9                 # the last allocation anywhere is rank 1's step 4 -> chunk [6-8] is quiet
    assert res[0][1] == res[1][1] == [2.0 * i + 1.0 for i in range(9)]
