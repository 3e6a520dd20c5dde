"""Per-shape timings of the tcgen05 weight-gradient kernel (k_wgrad_tc) at the bench's level sizes (GPU box only).

  [SCN_B200_WG_PAIRS=<pairs per CTA>] python tools/wgrad_sweep.py
One line per shape: us/launch (back-to-back launches behind a GPU-side delay), algorithmic TFLOP/s, gathered GB/s
(P x (Cin + Cout) x 2 B: both operands of a pair are gathered), and the error against an fp32 torch contraction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparseeventid_b200 import _lib as L
from sparseeventid_b200.scn import ops

dev = "cuda"
SHAPES = [(495518, 27, 32, 32, 0.19), (317485, 27, 64, 64, 0.32), (154605, 27, 96, 96, 0.44), (59700, 27, 128, 128, 0.48),
          (20727, 27, 160, 160, 0.44), (7332, 27, 192, 192, 0.36), (317485, 8, 32, 64, 0.195), (154605, 8, 64, 96, 0.26),
          (20727, 8, 128, 160, 0.36), (7332, 8, 160, 192, 0.35), (7332, 1, 192, 128, 1.0)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
for si, (n, K, cin, cout, dens) in enumerate(SHAPES):
    if only >= 0 and si != only:
        continue
    torch.manual_seed(0)
    n_pad = ops.pad128(n)
    nbr = torch.full((K, n_pad), -1, dtype=torch.int32, device=dev)
    base = torch.arange(n, device=dev, dtype=torch.int32)[None, :].expand(K, n)
    idx = (base + torch.randint(-40, 41, (K, n), device=dev, dtype=torch.int32)).clamp(0, n - 1)
    mask = torch.rand(K, n, device=dev) < dens
    nbr[:, :n] = torch.where(mask, idx, torch.full_like(idx, -1))
    if K % 2 == 1:
        nbr[K // 2, :n] = torch.arange(n, device=dev, dtype=torch.int32)
    x = torch.randn(n, cin, device=dev).bfloat16()
    d = torch.randn(n, cout, device=dev).bfloat16()
    dw = ops.conv_wgrad(x, d, nbr, n, cin, cout, L.PREC_BF16)
    torch.cuda.synchronize()
    ref = torch.zeros(K, cin, cout, device=dev)
    xf, df = x.float(), d.float()
    for k in range(K):
        j = nbr[k, :n].long()
        m = j >= 0
        ref[k] = xf[j[m]].t() @ df[m]
    err = float((dw - ref).norm() / ref.norm())
    out = torch.zeros((K, cin, cout), dtype=torch.float32, device=dev)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for _ in range(reps):
        L.check(L.lib().scn_conv_wgrad(L.ptr(x), 1, L.ptr(d), 1, L.ptr(nbr), K, n, n_pad, cin, cout, 1, L.ptr(out), L.stream()), "wgrad")
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    pairs = int((nbr >= 0).sum())
    print(f"n={n:7d} K={K:3d} {cin:3d}->{cout:3d} pairs {pairs:8d}: {us:7.1f} us  {2.0 * pairs * cin * cout / us / 1e6:6.1f} TF/s  "
          f"gathered {pairs * (cin + cout) * 2 / us / 1e3:6.0f} GB/s  err {err:.1e}{'  WRONG' if err > 5e-3 else ''}", flush=True)
