"""Shared test helpers (seeded random sparse inputs)."""
import numpy as np
import torch


def random_sites(n, grid, batch, seed, dup=0):
    """int64 [n(+dup), 4] rows (x, y, z, b), unique unless dup > 0; batch index last."""
    rng = np.random.default_rng(seed)
    g = np.asarray(grid)
    total = int(np.prod(g)) * batch
    flat = rng.choice(total, size=min(n, total), replace=False)
    b, r = np.divmod(flat, int(np.prod(g)))
    x, r = np.divmod(r, g[1] * g[2])
    y, z = np.divmod(r, g[2])
    c = np.stack([x, y, z, b], 1).astype(np.int64)
    if dup:
        c = np.concatenate([c, c[rng.integers(0, c.shape[0], size=dup)]], 0)
        c = c[rng.permutation(c.shape[0])]
    return c


def blob_sites(n, grid, batch, seed):
    """Clustered sites (random walk) so that neighbourhoods are dense like tracks."""
    rng = np.random.default_rng(seed)
    g = np.asarray(grid)
    out = []
    for b in range(batch):
        p = (g // 2).astype(np.int64)
        pts = []
        for _ in range(n):
            p = np.clip(p + rng.integers(-1, 2, size=3), 0, g - 1)
            pts.append(p.copy())
        pts = np.unique(np.asarray(pts), axis=0)
        pts = pts[rng.permutation(pts.shape[0])]
        out.append(np.concatenate([pts, np.full((pts.shape[0], 1), b)], 1))
    return np.concatenate(out, 0).astype(np.int64)


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
