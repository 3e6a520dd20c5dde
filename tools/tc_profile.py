"""One conv shape, a few launches: the command profiled by ncu (see profiles/).

  python tools/tc_profile.py <rows> <K> <Cin> <Cout> [reps]
The table is track-like random with the centre offset set to the identity (what a submanifold table has).
Timing: a GPU-side delay first so that the host is ahead of the GPU, then `reps` back-to-back launches between two
events (average per launch; L2 is not flushed: activations of these sizes are L2-resident in the network too)."""
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
x = torch.randn(n, cin, device=dev).bfloat16()
w = (torch.randn(K, cin, cout, device=dev) / cin ** 0.5).contiguous()
bp = ops.prep_weights(w, False, False, L.PREC_BF16, torch.bfloat16)
out = torch.empty((n, cout), dtype=torch.bfloat16, device=dev)


def launch():
    L.check(L.lib().scn_conv_forward(L.ptr(x), 1, n, L.ptr(nbr), K, n, n_pad, cin, cout, L.ptr(bp), None, 1, L.ptr(out), 1,
                                     L.stream()), "conv")


launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(4_000_000)
e0.record()
for i in range(reps):
    launch()
e1.record()
torch.cuda.synchronize()
pairs = int((nbr >= 0).sum())
us = e0.elapsed_time(e1) * 1e3 / reps
print(f"n={n} K={K} {cin}->{cout}: {us:.1f} us/launch over {reps}, pairs={pairs}, "
      f"{2.0 * pairs * cin * cout / us / 1e6:.1f} TFLOP/s algorithmic", flush=True)
if os.environ.get("TC_PROFILE_CHECK", "1") != "0":
    # fp32 reference of the same contraction on the bf16-rounded operands
    xf, wf = x.float(), w.bfloat16().float()
    ref = torch.zeros(n, cout, device=dev)
    for k in range(K):
        j = nbr[k, :n].long()
        live = j >= 0
        ref[live] += xf[j[live]] @ wf[k]
    err = float((out.float() - ref).norm() / ref.norm())
    print(f"  rel L2 error vs fp32 torch on the same bf16 operands: {err:.2e}", flush=True)
    assert err < 5e-3, "tc_profile: wrong result"
