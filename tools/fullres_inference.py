"""BASELINE.json configs[4] / SURVEY.md §8d config 5: full-resolution 3-D DUNE events (1536^3 grid, the legacy
torch/sparseresnet3d.ResNet shape: 5^3 stem, residual series + stride-2 downsample per level, nf 32, 2 blocks per
level, BatchNorm in eval mode) -- deep-ResNet INFERENCE events/s, and the batch that 180 GB of HBM would hold
(extrapolated from the measured peak memory per event; the packed site key carries the batch index in 16 bits,
so one InputLayer call takes at most 65535 events).

    python tools/fullres_inference.py [--batch 256] [--depth 9] [--steps 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sparseconvnet as scn
from sparseeventid_b200 import legacy_networks as legacy
from sparseeventid_b200 import synthetic
from sparseeventid_b200.data_transforms import larcvsparse_to_scnsparse_3d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--depth", type=int, default=9)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join("gpurun_out", "fullres_inference.json"))
    ap.add_argument("--bytes-per-event", type=float, default=0.0, help="measured at a smaller batch: enables the memory guard")
    a = ap.parse_args()
    scn.set_precision("bf16")
    cfg = legacy.LegacyNetworkConfig(n_initial_filters=32, network_depth=a.depth, res_blocks_per_layer=2,
                                     batch_norm=True, leaky_relu=False)
    torch.manual_seed(0)
    model = legacy.LegacyResNet3D(scn, legacy.LEGACY_OUTPUT_SHAPE, cfg).cuda().eval()
    arr = synthetic.larcv_batch_3d(a.batch, seed=9000, grid=(1536, 1536, 1536))
    c, f, b = larcvsparse_to_scnsparse_3d(arr)
    coords = torch.from_numpy(np.ascontiguousarray(c)).cuda()
    feats = torch.from_numpy(np.ascontiguousarray(f)).float().cuda()
    n = coords.shape[0]
    if a.bytes_per_event > 0:                 # memory guard for the large batches: never drive the box out of memory
        free, total = torch.cuda.mem_get_info()
        need = a.bytes_per_event * a.batch * 1.2
        if need > 0.85 * free:
            print(json.dumps({"batch": a.batch, "skipped": f"predicted {need / 1e9:.0f} GB > 85% of the free {free / 1e9:.0f} GB"}))
            return
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        for _ in range(2):
            out = model((coords, feats, b))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = model((coords, feats, b))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    peak = torch.cuda.max_memory_allocated() - base
    per_event = peak / a.batch
    res = {"workload": f"legacy sparseresnet3d shape, 1536^3 grid, depth {a.depth}, nf 32, 2 blocks/level, eval-mode BN, bf16",
           "batch": a.batch, "voxels": n, "voxels_per_event": n / a.batch, "ms_per_batch": ms,
           "events_per_s": a.batch / (ms * 1e-3), "peak_bytes_above_weights_and_input": peak,
           "bytes_per_event": per_event, "max_batch_180GB_extrapolated": int(170e9 / per_event),
           "hbm_peak_allocated_GB": torch.cuda.max_memory_allocated() / 1e9,
           "max_batch_per_call": 65535, "logits_finite": bool(all(torch.isfinite(v).all() for v in out.values()))}
    print(json.dumps(res))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "a") as fo:
        fo.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
