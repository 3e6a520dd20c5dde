"""larcv batch-filler array -> SparseConvNet input tuple (host side, numpy).

Mirror of the reference's input contract; same names, argument meaning and
output layout as ``src/io/data_transforms.py:21-49`` (3-D) and ``:198-252`` (2-D), written
without the reference's unconditional ``torch_geometric`` import (``:2``).
"""
from __future__ import annotations

import numpy

PAD_VALUE = -999


def larcvsparse_to_scnsparse_3d(input_array):
    """``[B, 1, MaxVoxels, 4]`` (x, y, z, value) -> (coords [N,4] = (x,y,z,batch), features [N,1], B).

    Rows keep numpy.where order (batch-major); coords carry the integral voxel indices in the
    array's floating dtype promoted with the int64 batch index, exactly like the reference.
    """
    value = input_array[..., -1]
    keep = numpy.where(value != PAD_VALUE)            # (batch, plane, voxel)
    batch_size = input_array.shape[0]
    features = numpy.expand_dims(value[keep], axis=-1)
    cols = [input_array[..., a][keep] for a in range(input_array.shape[-1] - 1)]
    cols.append(keep[0])
    coords = numpy.stack(cols, axis=-1)
    return (coords, features, batch_size,)


def larcvsparse_to_scnsparse_2d(input_array):
    """``[B, planes, MaxVoxels, 3]`` (x, y, value) -> [coords [N,4] = (plane,y,x,batch), features, B].

    Rows are plane-major: every sample's plane-0 voxels, then plane 1, then plane 2.
    """
    n_planes = input_array.shape[1]
    batch_size = input_array.shape[0]
    all_coords, all_features = [], []
    for p in range(n_planes):
        plane = input_array[:, p]
        value = plane[..., 2]
        keep = numpy.where(value != PAD_VALUE)
        x = plane[..., 0][keep]
        y = plane[..., 1][keep]
        pl = numpy.full(x.shape, fill_value=p)
        all_coords.append(numpy.stack([pl, y, x, keep[0]], axis=-1))
        all_features.append(numpy.expand_dims(value[keep], axis=-1))
    return [numpy.concatenate(all_coords), numpy.concatenate(all_features), batch_size]


# ------------------------------------------------------------------------------------------------
# Device-side twins: same outputs, same row order, no host pass.  The pinned larcv buffer is
# copied to the GPU as it is; the compaction runs in libscn_b200.so.  Coordinates come back as int32 [N, 4] on the
# device (scn.InputLayer takes any integer or floating dtype), features as fp32 [N, 1].
# ------------------------------------------------------------------------------------------------


def _larcv_to_scn_gpu(input_array, layout):
    import torch

    from . import _lib as L
    t = torch.as_tensor(input_array)
    L.require_cuda(t, "larcvsparse_to_scnsparse_gpu")
    t = t.contiguous().float()
    B, P, V, ncol = t.shape
    lib = L.lib()
    counts = torch.empty((P * B,), dtype=torch.int32, device=t.device)
    L.check(lib.scn_larcv_count(t.data_ptr(), B, P, V, ncol, float(PAD_VALUE), counts.data_ptr(), L.stream()),
            "scn_larcv_count")
    incl = torch.cumsum(counts, 0, dtype=torch.int64)
    offs = (incl - counts).contiguous()
    n = int(incl[-1].item())                       # the one host read-back: rows to allocate
    coords = torch.empty((n, 4), dtype=torch.int32, device=t.device)
    feats = torch.empty((n, 1), dtype=torch.float32, device=t.device)
    if n:
        L.check(lib.scn_larcv_compact(t.data_ptr(), B, P, V, ncol, float(PAD_VALUE), offs.data_ptr(), layout,
                                      coords.data_ptr(), feats.data_ptr(), L.stream()), "scn_larcv_compact")
    return coords, feats, B


def larcvsparse_to_scnsparse_3d_gpu(input_array):
    """Device twin of :func:`larcvsparse_to_scnsparse_3d`: CUDA tensor ``[B, 1, MaxVoxels, 4]`` -> (coords int32
    [N,4] = (x,y,z,batch), features [N,1], B), rows in the same (batch, voxel) order."""
    coords, feats, b = _larcv_to_scn_gpu(input_array, 0)
    return (coords, feats, b,)


def larcvsparse_to_scnsparse_2d_gpu(input_array):
    """Device twin of :func:`larcvsparse_to_scnsparse_2d`: ``[B, planes, MaxVoxels, 3]`` -> [coords int32 [N,4] =
    (plane,y,x,batch), features [N,1], B], rows plane-major like the reference."""
    coords, feats, b = _larcv_to_scn_gpu(input_array, 1)
    return [coords, feats, b]
