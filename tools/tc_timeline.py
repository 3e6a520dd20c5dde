"""Per-role clock64 timeline of CTA 0 of the tcgen05 conv kernel (debug; GPU box only).
    python tools/tc_timeline.py n K cin cout"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sparseeventid_b200 import _lib as L
from sparseeventid_b200.scn import ops

n, K, cin, cout = [int(v) for v in sys.argv[1:5]]
torch.manual_seed(0)
dev = "cuda"
n_pad = ops.pad128(n)
nbr = torch.full((K, n_pad), -1, dtype=torch.int32, device=dev)
base = torch.arange(n, device=dev, dtype=torch.int32)[None, :].expand(K, n)
idx = (base + torch.randint(-40, 41, (K, n), device=dev, dtype=torch.int32)).clamp(0, n - 1)
mask = torch.rand(K, n, device=dev) < 0.3
nbr[:, :n] = torch.where(mask, idx, torch.full_like(idx, -1))
x = torch.randn(n, cin, device=dev).bfloat16()
w = (torch.randn(K, cin, cout, device=dev) / cin ** 0.5).contiguous()
bp = ops.prep_weights(w, False, False, L.PREC_BF16, torch.bfloat16)
out = torch.empty((n, cout), dtype=torch.bfloat16, device=dev)
lib = L.lib()
lib.scn_tc_debug_timeline.argtypes = [ctypes.c_void_p]
lib.scn_tc_debug_timeline.restype = None
def run():
    L.check(lib.scn_conv_forward(L.ptr(x), 1, n, L.ptr(nbr), K, n, n_pad, cin, cout, L.ptr(bp), None, 1, L.ptr(out), 1, L.stream()), "conv")
run(); run()
dbg = torch.zeros(4 * 256 * 8, dtype=torch.int64, device=dev)
lib.scn_tc_debug_timeline(dbg.data_ptr())
run()
torch.cuda.synchronize()
lib.scn_tc_debug_timeline(None)
d = dbg.cpu().numpy().reshape(4, 256, 8)
t0 = d[d > 0].min()
d = np.where(d > 0, d - t0, -1)
print("producer stage: start list_done slot_free issued published   (cycles since first mark)")
for st in range(0, 64):
    print("P", st, *d[0, st, :5])
print("mma stage: poll_start ready issued")
for st in range(0, 64):
    print("M", st, *d[1, st, :3])
print("bload tile: wait_start slot_free published ; mma bfull wait start/end")
for st in range(0, 40):
    print("B", st, *d[2, st, :3], "|", *d[3, st, :2])
pub = d[0, :200, 4]; pub = pub[pub > 0]
print("median cycles between consecutive published stages:", np.median(np.diff(np.sort(pub))))
iss = d[1, :200, 2]; iss = iss[iss > 0]
print("median cycles between consecutive MMA issues:", np.median(np.diff(np.sort(iss))))
