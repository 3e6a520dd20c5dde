"""bench.py's voxel-count-balanced event sharding (N > 1): the ranks' shares partition the global batch, every rank gets
the same number of events and nearly the same number of voxels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_snake_deal_partitions_and_balances():
    import bench
    from sparseeventid_b200 import synthetic
    world, batch, seed = 4, 6, 321
    shares = [bench.balanced_host_batch(batch, world, r, seed, "dune3d") for r in range(world)]
    arr = synthetic.larcv_batch_3d(batch * world, seed=seed)
    total = int((arr[..., -1] != synthetic.PAD).sum())
    voxels = [s[0].shape[0] for s in shares]
    assert sum(voxels) == total                                   # nothing lost, nothing duplicated
    assert all(s[2] == batch and int(s[0][:, 3].max()) == batch - 1 for s in shares)
    assert all(len(v) == batch for s in shares for v in s[3].values())
    assert max(voxels) <= 1.08 * (total / world), voxels          # contiguous sharding of the same events: up to ~1.3x
    # the shares are distinct event sets: total feature sums add up to the global one
    fsum = sum(float(s[1].astype(np.float64).sum()) for s in shares)
    gsum = float(arr[..., -1][arr[..., -1] != synthetic.PAD].astype(np.float64).sum())
    assert abs(fsum - gsum) < 1e-6 * max(1.0, abs(gsum))
