"""World-size-2 gloo test (CPU, oracle modules) of the WHOLE event-sharded training step in sparseeventid_b200/trainer.py:
rank 0's initial weights are broadcast, every rank runs the sparse encoder + heads on ITS events, the flat gradient
arena is averaged across ranks, Adam steps -- after which all ranks must hold identical parameters, equal to a
single-process run that averages the two ranks' gradients by hand (what DDP / Horovod average does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

CFG = dict(depth=2, n_initial_filters=8, n_output_filters=8, blocks_per_layer=1)
IMAGE = (1, 32, 32, 32)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(rank):
    """Two tiny events per rank on a 32^3 grid (different on every rank)."""
    rng = np.random.default_rng(100 + rank)
    rows = []
    for b in range(2):
        pts = np.unique(rng.integers(8, 24, size=(60, 3)), axis=0)
        rows.append(np.concatenate([pts, np.full((pts.shape[0], 1), b)], 1))
    coords = np.concatenate(rows, 0).astype(np.int64)
    feats = rng.normal(size=(coords.shape[0], 1)).astype(np.float32)
    labels = {"labelneutID": np.array([0, 2]), "labelprotID": np.array([1, 0]), "labelnpiID": np.array([1, 0]),
              "labelcpiID": np.array([0, 1])}
    return (torch.as_tensor(coords), torch.as_tensor(feats), 2), {k: torch.as_tensor(v) for k, v in labels.items()}


def _make_trainer(seed):
    from oracle import sparseconvnet_oracle as oscn
    from sparseeventid_b200 import networks
    oscn.set_numerics("fp32")
    torch.manual_seed(seed)
    cfg = networks.EncoderConfig(**CFG)
    enc = networks.Encoder(oscn, cfg, IMAGE, 3)
    head = networks.MultiHeadOutput(enc.output_shape[1:], enc.output_shape[0], networks.OUTPUT_SHAPE)
    return networks, networks.EventIDModel(enc, head)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparseeventid_b200.trainer import FlatGradArena
    networks, model = _make_trainer(seed=rank)            # different initial weights: rank 0's must win
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=0)
    model.train()
    model.head.eval()                                     # Dropout off: the comparison is deterministic
    arena = FlatGradArena(list(model.parameters()))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, eps=1e-6, betas=(0.8, 0.9), weight_decay=1e-6)
    batch, labels = _batch(rank)
    for _ in range(2):
        arena.zero()
        loss = networks.focal_loss(labels, model(batch))
        loss.backward()
        arena.finish()
        opt.step()
    q.put((rank, [p.detach().numpy().copy() for p in model.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_event_sharded_training_step_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, params = q.get(timeout=300)
        res[r] = params
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)                       # identical parameters on both ranks

    # single process: two replicas (one per rank's events) whose gradients are averaged by hand
    networks, model = _make_trainer(seed=0)
    model.train()
    model.head.eval()
    _, twin = _make_trainer(seed=0)
    twin.train()
    twin.head.eval()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, eps=1e-6, betas=(0.8, 0.9), weight_decay=1e-6)
    b0, l0 = _batch(0)
    b1, l1 = _batch(1)
    for _ in range(2):
        twin.load_state_dict(model.state_dict())         # BatchNorm statistics are per rank: replica state is irrelevant here
        model.zero_grad()
        twin.zero_grad()
        networks.focal_loss(l0, model(b0)).backward()
        networks.focal_loss(l1, twin(b1)).backward()
        for p, t in zip(model.parameters(), twin.parameters()):
            p.grad = (p.grad + t.grad) / 2
        opt.step()
    for got, p in zip(res[0], model.parameters()):
        assert np.allclose(got, p.detach().numpy(), rtol=1e-5, atol=1e-6)
