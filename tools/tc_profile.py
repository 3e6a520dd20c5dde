"""One conv shape, a few launches: the command profiled by ncu (see profiles/).

  python tools/tc_profile.py <rows> <K> <Cin> <Cout> [reps] [lists]
`lists`: also run the experimental stage-list kernel (k_conv_tcl) on the same table, time it and compare the outputs.
The table is track-like random with the centre offset set to the identity (what a submanifold table has)."""
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
if K % 2 == 1:
    nbr[(K - 1) // 2, :n] = torch.arange(n, device=dev, dtype=torch.int32)      # submanifold: the centre offset is the identity
use_lists = len(sys.argv) > 6 and sys.argv[6] == "lists"
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
print(f"n={n} K={K} {cin}->{cout}: last launch {e0.elapsed_time(e1)*1e3:.1f} us, pairs={int((nbr >= 0).sum())}")
if use_lists:
    lists = ops.stage_lists(nbr)
    out2 = torch.empty_like(out)
    torch.cuda.synchronize()
    for i in range(reps):
        if i == reps - 1:
            e0.record()
        L.check(L.lib().scn_conv_forward_sl(L.ptr(x), 1, n, L.ptr(nbr), K, n, n_pad, cin, cout, L.ptr(bp), None, 1,
                                            L.ptr(out2), 1, L.ptr(lists), L.stream()), "conv_sl")
    e1.record()
    torch.cuda.synchronize()
    d = (out2.float() - out.float()).abs()
    ref = out.float().abs()
    print(f"  stage lists: last launch {e0.elapsed_time(e1)*1e3:.1f} us; vs default kernel: max |diff| {float(d.max()):.4f} "
          f"(max |out| {float(ref.max()):.2f}), rows off by more than rounding: {int((d > 0.02 * ref + 0.02).any(1).sum())}")
