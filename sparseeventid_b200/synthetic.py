"""Synthetic DUNE-shaped sparse events in larcv3 batch-filler layout (host side, numpy).

No dataset ships with the reference (README.md:19), so the workload is synthesised to the
shapes the reference fixes:
  * grid 1024x512x1280 for dune3d, 3 planes of 1536x1024 for dune2d (src/io/larcv_fetcher.py:23-48)
  * larcv BatchFillerSparseTensor output ``[B, planes, MaxVoxels=50000, D+1]`` float32 padded with
    -999 (src/io/larcv_fetcher.py:112-119; consumed by src/io/data_transforms.py:21-49,198-252)
  * voxel values normalised ``(q - 1.0) / 0.5`` (src/io/larcv_fetcher.py:100-108)
  * label keys / class counts (src/utils/supervised_eventID.py:224-229)
Generator = SURVEY.md App. D: per event a vertex in the central 60 % of the grid, 2-6 straight
tracks (length 100-700 voxels, 0.5-voxel steps, sigma 0.6 transverse smear, two samples per
step) and 0-2 showers (2000-8000 points, gamma depth, cone spread 0.05 t + 1).
"""
from __future__ import annotations

import numpy as np

GRID_3D = (1024, 512, 1280)
GRID_2D = (3, 1536, 1024)
MAX_VOXELS = 50000
PAD = -999.0
LABEL_CLASSES = {"labelneutID": 3, "labelprotID": 3, "labelnpiID": 2, "labelcpiID": 2}


def _unit(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def _event_points(rng, grid):
    g = np.asarray(grid, dtype=np.float64)
    vertex = g * (0.2 + 0.6 * rng.random(3))
    pts = []
    for _ in range(int(rng.integers(2, 7))):
        d = _unit(rng)
        length = rng.uniform(100, 700)
        t = np.arange(0.0, length, 0.5)
        t = np.repeat(t, 2)
        p = vertex[None, :] + t[:, None] * d[None, :] + rng.normal(scale=0.6, size=(t.shape[0], 3))
        pts.append(p)
    for _ in range(int(rng.integers(0, 3))):
        d = _unit(rng)
        n = int(rng.integers(2000, 8001))
        t = rng.gamma(shape=3.0, scale=40.0, size=n)
        spread = 0.05 * t + 1.0
        start = vertex + d * rng.uniform(0, 50)
        p = start[None, :] + t[:, None] * d[None, :] + rng.normal(size=(n, 3)) * spread[:, None]
        pts.append(p)
    p = np.floor(np.concatenate(pts, 0)).astype(np.int64)
    ok = np.all((p >= 0) & (p < np.asarray(grid)[None, :]), axis=1)
    p = p[ok]
    p = np.unique(p, axis=0)
    return p


def _values(rng, n):
    q = rng.lognormal(mean=0.0, sigma=0.5, size=n)
    return ((q - 1.0) / 0.5).astype(np.float32)


def make_labels(batch, seed=1234):
    out = {}
    for i, (k, c) in enumerate(LABEL_CLASSES.items()):
        rng = np.random.default_rng(seed * 7 + i)
        out[k] = rng.integers(0, c, size=batch).astype(np.int64)
    return out


def larcv_batch_3d(batch, seed=1234, grid=GRID_3D, max_voxels=MAX_VOXELS, first_event=0):
    """[B, 1, max_voxels, 4] float32 = (x, y, z, value), padded with -999."""
    out = np.full((batch, 1, max_voxels, 4), PAD, dtype=np.float32)
    for b in range(batch):
        rng = np.random.default_rng(seed + first_event + b)
        p = _event_points(rng, grid)
        if p.shape[0] > max_voxels:
            p = p[rng.permutation(p.shape[0])[:max_voxels]]
        n = p.shape[0]
        out[b, 0, :n, :3] = p
        out[b, 0, :n, 3] = _values(rng, n)
    return out


def larcv_batch_2d(batch, seed=1234, max_voxels=MAX_VOXELS, first_event=0):
    """[B, 3, max_voxels, 3] float32 = (x, y, value): three 2-D projections of one 3-D event."""
    grid3 = (1024, 1536, 1024)            # (x, y, u): plane p projects out one axis
    out = np.full((batch, 3, max_voxels, 3), PAD, dtype=np.float32)
    for b in range(batch):
        rng = np.random.default_rng(seed + first_event + b)
        p = _event_points(rng, grid3)
        proj = [
            np.stack([p[:, 0], p[:, 1]], 1),
            np.stack([p[:, 2], p[:, 1]], 1),
            np.stack([(p[:, 0] + p[:, 2]) // 2, p[:, 1]], 1),
        ]
        for pl in range(3):
            q = np.unique(proj[pl], axis=0)
            ok = (q[:, 0] < GRID_2D[2]) & (q[:, 1] < GRID_2D[1])
            q = q[ok]
            if q.shape[0] > max_voxels:
                q = q[rng.permutation(q.shape[0])[:max_voxels]]
            n = q.shape[0]
            out[b, pl, :n, :2] = q
            out[b, pl, :n, 2] = _values(rng, n)
    return out


def uniform_cube(n, occupancy, seed, batch=1):
    """Config 4 sweep inputs: n sites drawn without replacement from an L^3 cube,
    L = ceil((n / occupancy)^(1/3)).  Returns int64 [n, 4] (x, y, z, b) and L."""
    rng = np.random.default_rng(seed)
    L = int(np.ceil((n / occupancy) ** (1.0 / 3.0)))
    rows = []
    for b in range(batch):
        flat = rng.choice(L ** 3, size=n, replace=False)
        x, r = np.divmod(flat, L * L)
        y, z = np.divmod(r, L)
        rows.append(np.stack([x, y, z, np.full_like(x, b)], 1))
    return np.concatenate(rows, 0).astype(np.int64), L
