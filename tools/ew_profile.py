"""Achieved bandwidth of the bandwidth-bound kernels at the bench workload's level shapes: BatchNorm forward / backward,
AddTable + LeakyReLU, LeakyReLU backward and the dense SparseToDense kernels (which the bench's lazy dense view never
launches).  Every number is the average of 20 back-to-back calls behind a GPU-side delay (kernel time on the stream, no
host latency); tensors of levels 0-2 are cycled through a pool larger than L2 so that reads come from HBM.

    python tools/ew_profile.py            (also the command profiled by ncu for profiles/*ew*)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparseeventid_b200.scn import ops

PEAK = 6549.8          # MEASURED_PEAKS.json hbm copy GB/s
shapes = [(495518, 32), (317485, 64), (154605, 96), (59700, 128), (20727, 160), (7332, 192)]
dev = "cuda"


def timed(fn, reps=20):
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


print(f"{'kernel':22s} {'rows':>7s} {'C':>4s} {'us':>8s} {'algorithmic GB/s':>17s} {'of ' + str(PEAK):>10s}")
for n, c in shapes:
    by = n * c * 2
    pool = max(2, min(8, int(300e6 // by) + 1))          # distinct tensors cycled: > L2 in total for the big levels
    xs = [torch.randn(n, c, device=dev).bfloat16() for _ in range(pool)]
    ds = [torch.randn(n, c, device=dev).bfloat16() for _ in range(pool)]
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev)
    rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev)
    _, stats = ops.bn_forward(xs[0], g, b, rm, rv, True, 1e-4, 0.9, 0.333)
    rows = [
        ("bn_fwd (2 launches)", 3, lambda i: ops.bn_forward(xs[i % pool], g, b, rm, rv, True, 1e-4, 0.9, 0.333)),
        ("bn_bwd (2 launches)", 5, lambda i: ops.bn_backward(xs[i % pool], ds[i % pool], g, b, stats, True, 0.333, want_colsum=True)),
        ("add+leaky fwd", 3, lambda i: ops.add_forward(xs[i % pool], ds[i % pool], 0.333)),
        ("leaky bwd", 3, lambda i: ops.leaky_backward(xs[i % pool], ds[i % pool], 0.333)),
    ]
    for name, passes, fn in rows:
        us = timed(fn)
        gbs = passes * by / us / 1e3
        print(f"{name:22s} {n:7d} {c:4d} {us:8.1f} {gbs:17.0f} {gbs / PEAK:10.2f}", flush=True)
    del xs, ds
    torch.cuda.empty_cache()

# SparseToDense at the bench's final level: 64 events, 128 channels after the bottleneck, 32 x 16 x 40 grid
B, C, sp, n = 64, 128, (32, 16, 40), 7332
gen = torch.Generator(device=dev).manual_seed(0)
cells = torch.randperm(B * sp[0] * sp[1] * sp[2], device=dev, generator=gen)[:n].sort().values
bi = cells // (sp[0] * sp[1] * sp[2]); r = cells % (sp[0] * sp[1] * sp[2])
x0, x1, x2 = r // (sp[1] * sp[2]), (r // sp[2]) % sp[1], r % sp[2]
keys = ((bi << 48) | (x0 << 32) | (x1 << 16) | x2).to(torch.int64)
x = torch.randn(n, C, device=dev).bfloat16()
dense = ops.sparse_to_dense_forward(x, keys, B, sp)
dd = torch.randn_like(dense)
us = timed(lambda i: ops.sparse_to_dense_forward(x, keys, B, sp), 10)
by = dense.numel() * 4 + x.numel() * 2
print(f"{'sparse_to_dense fwd':22s} {n:7d} {C:4d} {us:8.1f} {by / us / 1e3:17.0f} {by / us / 1e3 / PEAK:10.2f}   (dense {dense.numel() * 4 / 1e6:.0f} MB written)")
us = timed(lambda i: ops.sparse_to_dense_backward(dd, keys, n, C, B, sp, torch.bfloat16), 10)
by = n * C * 4 + n * C * 2 + n * 8
print(f"{'sparse_to_dense bwd':22s} {n:7d} {C:4d} {us:8.1f} {by / us / 1e3:17.0f} {by / us / 1e3 / PEAK:10.2f}   (reads the {n} active columns only)")
