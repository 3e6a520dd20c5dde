"""Per-shape timings of the tcgen05 convolution (k_conv_tc) at the bench's level sizes, with knob sweeps.

  python tools/tc_sweep.py [quick]
Prints one line per (shape, knobs): us/launch (average of back-to-back launches behind a GPU-side delay) and the
algorithmic TFLOP/s; every configuration's output is checked against an fp32 torch contraction of the same bf16 operands."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparseeventid_b200 import _lib as L
from sparseeventid_b200.scn import ops

lib = L.lib()
lib.scn_tc_debug_knobs.argtypes = [ctypes.c_int] * 4
lib.scn_tc_debug_knobs.restype = None
dev = "cuda"


def make(n, K, cin, cout, seed=0):
    torch.manual_seed(seed)
    n_pad = ops.pad128(n)
    nbr = torch.full((K, n_pad), -1, dtype=torch.int32, device=dev)
    base = torch.arange(n, device=dev, dtype=torch.int32)[None, :].expand(K, n)
    idx = (base + torch.randint(-40, 41, (K, n), device=dev, dtype=torch.int32)).clamp(0, n - 1)
    mask = torch.rand(K, n, device=dev) < 0.3
    nbr[:, :n] = torch.where(mask, idx, torch.full_like(idx, -1))
    if K % 2 == 1:
        nbr[(K - 1) // 2, :n] = torch.arange(n, device=dev, dtype=torch.int32)
    x = torch.randn(n, cin, device=dev).bfloat16()
    w = (torch.randn(K, cin, cout, device=dev) / cin ** 0.5).contiguous()
    bp = ops.prep_weights(w, False, False, L.PREC_BF16, torch.bfloat16)
    xf, wf = x.float(), w.bfloat16().float()
    ref = torch.zeros(n, cout, device=dev)
    for k in range(K):
        j = nbr[k, :n].long()
        live = j >= 0
        ref[live] += xf[j[live]] @ wf[k]
    return dict(n=n, K=K, cin=cin, cout=cout, n_pad=n_pad, nbr=nbr, x=x, bp=bp, ref=ref, pairs=int((nbr >= 0).sum()))


def run(c, reps=10, grid=0, t=0, sa=0, sb=0):
    lib.scn_tc_debug_knobs(grid, t, sa, sb)
    out = torch.empty((c["n"], c["cout"]), dtype=torch.bfloat16, device=dev)

    def launch():
        L.check(lib.scn_conv_forward(L.ptr(c["x"]), 1, c["n"], L.ptr(c["nbr"]), c["K"], c["n"], c["n_pad"], c["cin"], c["cout"],
                                     L.ptr(c["bp"]), None, 1, L.ptr(out), 1, L.stream()), "conv")
    launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    err = float((out.float() - c["ref"]).norm() / c["ref"].norm())
    tag = f"grid={grid or 'auto'} T={t or 'auto'} SA={sa or 'auto'} SB={sb or 'auto'}"
    print(f"n={c['n']:7d} K={c['K']:3d} {c['cin']:3d}->{c['cout']:3d} [{tag}]: {us:7.1f} us  "
          f"{2.0 * c['pairs'] * c['cin'] * c['cout'] / us / 1e6:6.1f} TF/s  err {err:.1e}{'  WRONG' if err > 5e-3 else ''}", flush=True)
    lib.scn_tc_debug_knobs(0, 0, 0, 0)
    return us


SHAPES = [(495518, 27, 32, 32), (317485, 27, 64, 64), (154605, 27, 96, 96), (59700, 27, 128, 128),
          (20727, 27, 160, 160), (7332, 27, 192, 192), (317485, 8, 32, 64), (154605, 8, 64, 96), (317485, 8, 64, 32),
          (7332, 1, 192, 128)]
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1          # one shape per process: a launch failure is sticky
for si, shp in enumerate(SHAPES):
    if only >= 0 and si != only:
        continue
    c = make(*shp)
    run(c)
    if quick:
        continue
    n, K, cin, cout = shp
    if K != 27:
        continue
    tmax = min(8, 512 // cout)
    for t in sorted({1, 2, 256 // cout, tmax} - {0}):
        run(c, t=t)
    for sa in (4, 8, 12):
        run(c, sa=sa)
    for sb in (2, 3, 4):
        run(c, sb=sb)
    if n < 30000:
        for grid in (74, 111):
            run(c, grid=grid)
    del c
    torch.cuda.empty_cache()
