"""One conv shape, a few launches: the command profiled by ncu (see profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparseeventid_b200 import _lib as L
from sparseeventid_b200.scn import ops

n, K, cin, cout = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (70000, 27, 64, 64))]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.manual_seed(0)
dev = "cuda"
n_pad = ops.pad128(n)
nbr = torch.full((K, n_pad), -1, dtype=torch.int32, device=dev)
# track-like locality: neighbours are nearby rows
base = torch.arange(n, device=dev, dtype=torch.int32)[None, :].expand(K, n)
idx = (base + torch.randint(-40, 41, (K, n), device=dev, dtype=torch.int32)).clamp(0, n - 1)
mask = torch.rand(K, n, device=dev) < 0.3
nbr[:, :n] = torch.where(mask, idx, torch.full_like(idx, -1))
x = torch.randn(n, cin, device=dev).bfloat16()
w = (torch.randn(K, cin, cout, device=dev) / cin ** 0.5).contiguous()
bp = ops.prep_weights(w, False, False, L.PREC_BF16, torch.bfloat16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = torch.empty((n, cout), dtype=torch.bfloat16, device=dev)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    L.check(L.lib().scn_conv_forward(L.ptr(x), 1, n, L.ptr(nbr), K, n, n_pad, cin, cout, L.ptr(bp), None, 1, L.ptr(out), 1,
                                     L.stream()), "conv")
e1.record()
torch.cuda.synchronize()
print(f"n={n} K={K} {cin}->{cout}: last launch {e0.elapsed_time(e1)*1e3:.1f} us, pairs={int(mask.sum())}")
