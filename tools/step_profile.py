"""Where does a training step go?  torch.profiler over a few steps of the bench workload:
GPU-busy time vs wall time (launch/CPU-bound share), top kernels, top CPU ops.

    python tools/step_profile.py [batch] [dataset] > gpurun_out/step_profile.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time

import torch
from torch.profiler import ProfilerActivity, profile

import sparseconvnet as scn
from bench import host_batch
from sparseeventid_b200.trainer import Trainer

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dataset = sys.argv[2] if len(sys.argv) > 2 else "dune3d"
dev = torch.device("cuda", 0)
scn.set_precision("bf16")
tr = Trainer(scn, dataset, device=dev, seed=0)
pool = []
for i in range(2):
    c, f, bs, lab = host_batch(batch, 1234 + 1000 * i, dataset)
    pool.append(((torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev), bs),
                 {k: torch.from_numpy(v).to(dev) for k, v in lab.items()}))
for i in range(3):
    tr.step(*pool[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(4):
    tr.step(*pool[i % 2])
t_cpu = time.perf_counter() - t0          # host time to ENQUEUE 4 steps (syncs inside the step included)
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"4 steps: host enqueue {t_cpu * 250:.2f} ms/step, wall {t_all * 250:.2f} ms/step")

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        tr.step(*pool[i % 2])
    torch.cuda.synchronize()
ka = prof.key_averages()
gpu_us = sum(e.self_device_time_total for e in ka)
print(f"GPU busy {gpu_us / 2e3:.2f} ms/step (sum of kernel+memcpy durations, 2 profiled steps)")
print(ka.table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=70))
print(ka.table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=70))
