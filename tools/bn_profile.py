"""Standalone timing of the BatchNorm kernels at the bench workload's level shapes (and the command profiled by ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparseeventid_b200.scn import ops

shapes = [(495518, 32), (317485, 64), (154605, 96), (59700, 128), (20727, 160), (7332, 192)]
if len(sys.argv) > 1:          # "3" = the first three shapes, "0,4" = shapes 0 and 4
    a = sys.argv[1]
    shapes = [shapes[int(i)] for i in a.split(",")] if "," in a else shapes[: int(a)]
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
        if not cold:
            # pipelined: 20 back-to-back calls behind a GPU-side delay (no host or event latency in the number; the
            # tensor stays L2-resident when it fits, as it is in the network right after the convolution wrote it)
            for name in ("fwd", "bwd"):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(2_000_000)
                e0.record()
                for _ in range(20):
                    if name == "fwd":
                        ops.bn_forward(x, g, b, rm, rv, True, 1e-4, 0.9, 0.333)
                    else:
                        ops.bn_backward(x, d, g, b, stats, True, 0.333)
                e1.record()
                torch.cuda.synchronize()
                ts[name + "_pipe"] = e0.elapsed_time(e1) * 1e3 / 20
            print(f"n={n} C={c} pipelined x20: fwd {ts['fwd_pipe']:7.1f} us  bwd {ts['bwd_pipe']:7.1f} us", flush=True)
        by = n * c * 2
        f, w = sorted(ts["fwd"])[2], sorted(ts["bwd"])[2]
        print(f"n={n} C={c} {'L2 flushed' if cold else 'warm      '}: fwd {f:7.1f} us ({3 * by / f / 1e3:6.0f} GB/s of 3NC)  "
              f"bwd {w:7.1f} us ({5 * by / w / 1e3:6.0f} GB/s of 5NC)", flush=True)
