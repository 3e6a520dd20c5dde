"""Per-role clock64 timeline of CTA 0 of the tcgen05 conv kernel + timing experiments (debug; GPU box only).

Needs the instrumented library:  SCN_B200_BUILD_TAG=dbg SCN_B200_NVCC_FLAGS=-DSCN_TC_TIMELINE python -m sparseeventid_b200.build
    SCN_B200_LIB=sparseeventid_b200/lib/libscn_b200_dbg.so python tools/tc_timeline.py n K cin cout [T]
Prints, for CTA 0: per producer group and per issuing warp the median cycles of each phase of a stage, the stage period,
and the launch time with parts of the kernel switched off (experiments: WRONG results, timing only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sparseeventid_b200 import _lib as L
from sparseeventid_b200.scn import ops

n, K, cin, cout = [int(v) for v in sys.argv[1:5]]
T = int(sys.argv[5]) if len(sys.argv) > 5 else 0
torch.manual_seed(0)
dev = "cuda"
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
out = torch.empty((n, cout), dtype=torch.bfloat16, device=dev)
lib = L.lib()
lib.scn_tc_debug_timeline.argtypes = [ctypes.c_void_p]
lib.scn_tc_debug_timeline.restype = None
lib.scn_tc_debug_exp.argtypes = [ctypes.c_int]
lib.scn_tc_debug_exp.restype = None
lib.scn_tc_debug_knobs.argtypes = [ctypes.c_int] * 4
lib.scn_tc_debug_knobs.restype = None
lib.scn_tc_debug_knobs(0, T, 0, 0)


def run():
    L.check(lib.scn_conv_forward(L.ptr(x), 1, n, L.ptr(nbr), K, n, n_pad, cin, cout, L.ptr(bp), None, 1, L.ptr(out), 1, L.stream()), "conv")


def timed(reps=10):
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


print(f"n={n} K={K} {cin}->{cout} T={T or 'auto'}")
for flags, what in ((0, "full kernel (instrumented build)"), (1, "no gather copies"), (2, "no MMAs"), (4, "no output stores"),
                    (3, "no gathers, no MMAs"), (7, "no gathers, no MMAs, no stores")):
    lib.scn_tc_debug_exp(flags)
    print(f"  exp {flags}: {timed():7.1f} us/launch   {what}")
lib.scn_tc_debug_exp(0)

ROLES = 16
dbg = torch.zeros(ROLES * 256 * 8, dtype=torch.int64, device=dev)
run(); run()
lib.scn_tc_debug_timeline(dbg.data_ptr())
run()
torch.cuda.synchronize()
lib.scn_tc_debug_timeline(None)
d = dbg.cpu().numpy().reshape(ROLES, 256, 8).astype(np.int64)
t0 = d[d > 0].min()
tend = d.max()
print(f"CTA 0 span: {tend - t0} cycles")


def phases(role, evs, names, skip=4):
    r = d[role]
    ok = (r[:, evs] > 0).all(1)
    ok[:skip] = False
    if ok.sum() < 3:
        return None
    r = r[ok]
    parts = [np.median(r[:, evs[i + 1]] - r[:, evs[i]]) for i in range(len(evs) - 1)]
    period = np.median(np.diff(r[:, evs[0]]))
    return f"{ok.sum():4d} stages  period {period:7.0f}  " + "  ".join(f"{nm} {v:6.0f}" for nm, v in zip(names, parts))


for g in range(4):
    s = phases(g, [0, 1, 2, 3], ["list+prefetch", "wait slot", "issue copies"])
    if s:
        print(f"producer group {g}: {s}")
for m in range(4):
    s = phases(4 + m, [0, 1, 2], ["wait landed", "mask+issue"])
    if s:
        print(f"issuing warp {m}  : {s}")
    s = phases(10 + m, [0, 1], ["wait weights"])
    if s:
        print(f"   per q          : {s}")
s = phases(8, [0, 1], ["wait slot"])
if s:
    print(f"weight loader   : {s}")
s = phases(9, [0, 1, 2], ["wait accumulators", "drain + store"], skip=0)
if s:
    print(f"epilogue warp 0 : {s}")
# raw view of a few steady-state stages of class 0
print("class 0, stages 8..15: producer (top, listed, slot free, issued) | issuer (wait start, landed, issued)   [cycles since first mark]")
for st in range(8, 16):
    a = [int(v - t0) if v > 0 else -1 for v in d[0, st, :4]]
    b = [int(v - t0) if v > 0 else -1 for v in d[4, st, :3]]
    print("  ", st, a, "|", b)
