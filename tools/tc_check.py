"""Diagnostic (GPU box only): tcgen05 conv kernel vs a plain torch gather+matmul of the same bf16 operands.
Prints error structure (by row%8, by 8-column chunk) so a descriptor/swizzle mistake is identifiable."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sparseeventid_b200 import _lib as L
from sparseeventid_b200.scn import ops

torch.manual_seed(0)
dev = "cuda"


def case(n, K, cin, cout, density=0.3, bias=True):
    n_pad = ops.pad128(n)
    nbr = torch.full((K, n_pad), -1, dtype=torch.int32, device=dev)
    mask = torch.rand(K, n, device=dev) < density
    idx = torch.randint(0, n, (K, n), device=dev, dtype=torch.int32)
    nbr[:, :n] = torch.where(mask, idx, torch.full_like(idx, -1))
    x = torch.randn(n, cin, device=dev).bfloat16()
    w = (torch.randn(K, cin, cout, device=dev) / (cin ** 0.5)).bfloat16().float()
    b = torch.randn(cout, device=dev) if bias else None
    path = ops.conv_path(K, cin, cout, L.PREC_BF16, torch.bfloat16)
    bp = ops.prep_weights(w.contiguous(), False, False, L.PREC_BF16, torch.bfloat16)
    torch.cuda.synchronize()
    out = ops.conv_forward(x, nbr, n, cin, cout, bp, b, L.PREC_BF16, torch.bfloat16)
    torch.cuda.synchronize()
    ref = torch.zeros(n, cout, device=dev)
    if b is not None:
        ref += b
    xf = x.float()
    for k in range(K):
        j = nbr[k, :n].long()
        m = j >= 0
        ref[m] += xf[j[m]] @ w[k]
    err = (out.float() - ref).abs()
    rel = float(err.max() / ref.abs().max())
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        ops.conv_forward(x, nbr, n, cin, cout, bp, b, L.PREC_BF16, torch.bfloat16)
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / 5 * 1e3
    pairs = int(mask.sum())
    print(f"n={n} K={K} {cin}->{cout} path={path} rel_max_err={rel:.3e} time={us:.1f}us "
          f"TF={2.0*pairs*cin*cout/us/1e6:.1f}", flush=True)
    if rel > 1e-2:
        e = err.cpu().numpy()
        print("  err by row%8:", [float(e[i::8].max()) for i in range(8)])
        print("  err by col chunk of 8:", [float(e[:, c:c + 8].max()) for c in range(0, cout, 8)])
        print("  err by row tile of 128 (first 8):", [float(e[t * 128:(t + 1) * 128].max()) for t in range(min(8, (n + 127) // 128))])
        print("  out[0,:8]", out[0, :8].float().tolist(), "ref[0,:8]", ref[0, :8].tolist())
    return rel


if __name__ == "__main__":
    bad = 0
    for (n, K, ci, co) in [(128, 1, 64, 64), (300, 3, 64, 64), (1000, 27, 32, 32), (5000, 27, 64, 64),
                           (3000, 27, 96, 96), (3000, 27, 128, 128), (2000, 27, 160, 160), (2000, 27, 192, 192),
                           (4000, 8, 32, 64), (4000, 8, 160, 192), (70000, 27, 64, 64), (1000, 1, 192, 128)]:
        bad += case(n, K, ci, co) > 1e-2
    print("TC_CHECK", "FAIL" if bad else "OK")
