"""BASELINE.json configs[3] / SURVEY.md §8d config 4: isolated SubmanifoldConvolution 3x3x3 sweep --
channels x active sites x occupancy, forward / dgrad / wgrad separately, against the roofline.

    python tools/conv_sweep.py [--full] [--out gpurun_out/conv_sweep.json]

Coordinates: n sites drawn without replacement from an L^3 cube, L = ceil((n/occ)^(1/3)) (numpy default_rng seed
C*1000 + log10(n)*10 + occ_idx up to 1e6 sites; drawn on the device above that), batch 1; features N(0,1),
weights N(0, sqrt(2/(27 C))).  Plus one "track-like" case per channel count from the App. D event generator
(P/N ~ 8, like the bench workload).  FLOPs = 2 P Cin Cout; compulsory bytes as SURVEY §8d; the bound printed is
min(tensor peak, arithmetic intensity x HBM peak), both peaks from MEASURED_PEAKS.json.
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sparseconvnet as scn
from bench import load_peaks
from sparseeventid_b200 import _lib as L
from sparseeventid_b200 import synthetic
from sparseeventid_b200.scn import ops


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def cube_coords(n, occ, seed):
    if n <= 1_000_000:
        c, Lc = synthetic.uniform_cube(n, occ, seed)
        return torch.from_numpy(c).cuda(), Lc
    Lc = int(math.ceil((n / occ) ** (1.0 / 3.0)))
    g = torch.Generator(device="cuda").manual_seed(seed)
    flat = torch.unique(torch.randint(0, Lc ** 3, (int(n * 1.02) + 1024,), device="cuda", generator=g))
    flat = flat[torch.randperm(flat.numel(), device="cuda", generator=g)[:n]]
    x, r = flat // (Lc * Lc), flat % (Lc * Lc)
    return torch.stack([x, r // Lc, r % Lc, torch.zeros_like(x)], 1), Lc


def one_case(coords, grid, C, peaks, label):
    n = coords.shape[0]
    x = scn.InputLayer(3, list(grid))((coords, torch.zeros(n, 1, device="cuda"), 1))
    nbr = x.metadata.subm_table(tuple(grid), (3, 3, 3))
    n_act = x.metadata.levels[tuple(grid)].n
    P = int((nbr[:, :n_act] >= 0).sum())
    torch.manual_seed(0)
    feats = torch.randn(n_act, C, device="cuda").bfloat16()
    dout = torch.randn(n_act, C, device="cuda").bfloat16()
    w = (torch.randn(27, C, C, device="cuda") * math.sqrt(2.0 / (27 * C))).contiguous()
    prec, dt = L.PREC_BF16, torch.bfloat16
    bf = ops.prep_weights(w, False, False, prec, dt)
    bt = ops.prep_weights(w, True, True, prec, dt)
    path = ops.conv_path(27, C, C, prec, dt)
    t_f = timed(lambda: ops.conv_forward(feats, nbr, n_act, C, C, bf, None, prec, dt))
    t_d = timed(lambda: ops.conv_forward(dout, nbr, n_act, C, C, bt, None, prec, dt))
    t_w = timed(lambda: ops.conv_wgrad(feats, dout, nbr, n_act, C, C, prec))
    flops = 2.0 * P * C * C
    by_f = n_act * C * 2 * 2 + 27 * C * C * 2 + 8 * P
    by_w = n_act * C * 2 * 2 + 27 * C * C * 4 + 8 * P
    tpeak, hpeak = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
    rec = {"case": label, "C": C, "sites": n_act, "pairs": P, "pairs_per_site": P / n_act, "kernel_path": path}
    for name, t, by in (("fwd", t_f, by_f), ("dgrad", t_d, by_f), ("wgrad", t_w, by_w)):
        bound = min(tpeak, flops / by * hpeak)
        rec[name] = {"us": t * 1e6, "tflops": flops / t / 1e12, "frac_tensor_peak": flops / t / tpeak,
                     "algorithmic_gbs": by / t / 1e9, "frac_hbm_peak": by / t / hpeak,
                     "roofline_bound_tflops": bound / 1e12, "frac_of_bound": flops / t / bound}
    print(f"{label:34s} C={C:3d} N={n_act:8d} P/N={P / n_act:5.2f} path={path} "
          f"fwd {t_f * 1e6:8.1f} us {flops / t_f / 1e12:6.1f} TF | dgrad {t_d * 1e6:8.1f} us | wgrad {t_w * 1e6:8.1f} us "
          f"{flops / t_w / 1e12:6.1f} TF | bound {min(tpeak, flops / by_f * hpeak) / 1e12:6.1f} TF", flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="the whole SURVEY grid (C 16..256, N up to 1e7)")
    ap.add_argument("--tracks-only", action="store_true", help="only the track-like case of every channel count")
    ap.add_argument("--out", default=os.path.join("gpurun_out", "conv_sweep.json"))
    a = ap.parse_args()
    scn.set_precision("bf16")
    peaks = load_peaks()
    chans = [16, 32, 64, 128, 256] if a.full else ([32, 64, 96, 128, 160, 192] if a.tracks_only else [32, 64, 128, 256])
    sizes = [10_000, 100_000, 1_000_000, 10_000_000] if a.full else [10_000, 100_000, 1_000_000]
    occs = [0.001, 0.01, 0.05]
    out = []
    for C in chans:
        for n in ([] if a.tracks_only else sizes):
            if n * C > 3_000_000_000 // 2:
                continue
            for oi, occ in enumerate(occs):
                coords, Lc = cube_coords(n, occ, C * 1000 + int(round(math.log10(n))) * 10 + oi)
                out.append(one_case(coords, (Lc, Lc, Lc), C, peaks, f"uniform occ={occ} L={Lc}"))
                del coords
                torch.cuda.empty_cache()
        # track-like neighbourhoods (the bench workload's generator): 32 DUNE-shaped events
        from sparseeventid_b200.data_transforms import larcvsparse_to_scnsparse_3d
        c, f, b = larcvsparse_to_scnsparse_3d(synthetic.larcv_batch_3d(32, seed=500 + C))
        coords = torch.from_numpy(np.ascontiguousarray(c)).cuda()
        coords[:, 3] = 0                      # one sample: the sweep is batch 1 (tracks of 32 events overlaid)
        coords = torch.unique(coords.long(), dim=0)
        out.append(one_case(coords, synthetic.GRID_3D, C, peaks, "track-like (App. D generator)"))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"peaks": {k: v for k, v in peaks.items() if not isinstance(v, dict)}, "cases": out}, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
