"""Standalone timing of the BatchNorm kernels at the bench workload's level shapes (and the command profiled by ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparseeventid_b200.scn import ops

shapes = [(495518, 32), (317485, 64), (154605, 96), (59700, 128), (20727, 160), (7332, 192)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n, c in shapes:
    x = torch.randn(n, c, device="cuda").bfloat16()
    d = torch.randn(n, c, device="cuda").bfloat16()
    g = torch.ones(c, device="cuda"); b = torch.zeros(c, device="cuda")
    rm = torch.zeros(c, device="cuda"); rv = torch.ones(c, device="cuda")
    res = {}
    for name, fn in (("fwd", lambda: ops.bn_forward(x, g, b, rm, rv, True, 1e-4, 0.9, 0.333)),):
        pass
    out, stats = ops.bn_forward(x, g, b, rm, rv, True, 1e-4, 0.9, 0.333)
    for cold in (False, True):
        ts = {"fwd": [], "bwd": []}
        for _ in range(5):
            for name in ("fwd", "bwd"):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if name == "fwd":
                    ops.bn_forward(x, g, b, rm, rv, True, 1e-4, 0.9, 0.333)
                else:
                    ops.bn_backward(x, d, g, b, stats, True, 0.333)
                e1.record()
                torch.cuda.synchronize()
                ts[name].append(e0.elapsed_time(e1) * 1e3)
        by = n * c * 2
        f, w = sorted(ts["fwd"])[2], sorted(ts["bwd"])[2]
        print(f"n={n} C={c} {'L2 flushed' if cold else 'warm      '}: fwd {f:7.1f} us ({3 * by / f / 1e3:6.0f} GB/s of 3NC)  "
              f"bwd {w:7.1f} us ({5 * by / w / 1e3:6.0f} GB/s of 5NC)", flush=True)
