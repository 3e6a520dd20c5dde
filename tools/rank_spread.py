"""Why does an 8-rank step run at 0.93 of 8x the single-GPU rate when the collective costs 1%?  On ONE GPU, times the
training step of every rank's batch of a simulated W-rank job (the batches bench.py --gpus W would deal), for both
dealing rules, and prints per pool index the mean and the max over ranks: a synchronous step runs at the max.

    python tools/rank_spread.py [W=8] [pool=2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import sparseconvnet as scn
from sparseeventid_b200 import synthetic
from sparseeventid_b200.data_transforms import larcvsparse_to_scnsparse_3d
from sparseeventid_b200.trainer import Trainer

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
POOL = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = 64
dev = torch.device("cuda", 0)
scn.set_precision("bf16")
tr = Trainer(scn, "dune3d", device=dev, seed=0)


def deal(order, rank):
    mine = []
    for j, e in enumerate(order):
        r = j % (2 * W)
        r = r if r < W else 2 * W - 1 - r
        if r == rank:
            mine.append(int(e))
    return np.sort(np.asarray(mine))


def step_ms(arr, mine, labels_all):
    c, f, bs = larcvsparse_to_scnsparse_3d(np.ascontiguousarray(arr[mine]))
    batch = (torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64)).to(dev),
             torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32)).to(dev), bs)
    lab = {k: torch.from_numpy(v[mine]).to(dev) for k, v in labels_all.items()}
    for _ in range(3):
        tr.step(batch, lab)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr.step(batch, lab)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5, c.shape[0]


for i in range(POOL):
    seed = 1234 + 1000 * i
    arr = synthetic.larcv_batch_3d(B * W, seed=seed)
    labels_all = synthetic.make_labels(B * W, seed=seed)
    vox = (arr[..., -1] != synthetic.PAD).sum(axis=(1, 2))
    cost = bench.event_cost(arr, "dune3d")
    for name, key in (("voxel-count", vox), ("cost-proxy", cost)):
        order = np.argsort(-key, kind="stable")
        res = [step_ms(arr, deal(order, r), labels_all) for r in range(W)]
        ms = np.asarray([r[0] for r in res])
        nv = np.asarray([r[1] for r in res])
        print(f"pool {i} {name:12s}: step ms per rank {np.round(ms, 2).tolist()}  mean {ms.mean():.2f} max {ms.max():.2f} "
              f"(+{100 * (ms.max() / ms.mean() - 1):.1f}%)  voxels max/mean {nv.max() / nv.mean():.3f}", flush=True)
