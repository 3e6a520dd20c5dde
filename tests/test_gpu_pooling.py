"""GPU parity of scn.AveragePooling (the reference's non-default ``Pooling`` down-sampling branch,
src/networks/sparse_building_blocks.py:150-154) against the CPU oracle, forward and backward, through the C ABI
(``scn_pool_rows``).  Same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from helpers import blob_sites, random_sites
from oracle import scn_oracle as O
from oracle import sparseconvnet_oracle as oscn
from test_gpu_parity import MODES, run_pair, scn  # noqa: F401  (scn is the module fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("pool", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("c,drop", [(32, 0), (5, 0), (36, 4), (7, 2)])
def test_average_pooling(scn, mode, pool, c, drop):
    grid, B = (16, 16, 16), 3
    coords = blob_sites(400, grid, B, seed=21)
    run_pair(scn, mode,
             lambda: [scn.AveragePooling(3, list(pool), list(pool), drop)],
             lambda: [oscn.AveragePooling(3, list(pool), list(pool), drop)],
             coords, grid, B, c)


@pytest.mark.parametrize("mode", MODES)
def test_pooling_then_convolutions(scn, mode):
    """The pooled level is a first-class grid: submanifold and strided convolutions run on it."""
    grid, B = (16, 16, 16), 2
    coords = random_sites(600, grid, B, seed=22)
    run_pair(scn, mode,
             lambda: [scn.AveragePooling(3, 2, 2), scn.SubmanifoldConvolution(3, 32, 64, 3, False),
                      scn.Convolution(3, 64, 32, 2, 2, False)],
             lambda: [oscn.AveragePooling(3, 2, 2), oscn.SubmanifoldConvolution(3, 32, 64, 3, False),
                      oscn.Convolution(3, 64, 32, 2, 2, False)],
             coords, grid, B, 32)


def test_average_pooling_rows_match_convolution(scn):
    """Output sites / row order are those of a Convolution with the same window."""
    grid, B = (12, 8, 20), 2
    coords = random_sites(300, grid, B, seed=23)
    f = torch.ones(coords.shape[0], 3).cuda()
    x = scn.InputLayer(3, list(grid))((torch.as_tensor(coords).cuda(), f, B))
    p = scn.AveragePooling(3, 2, 2)(x)
    out_coords, rules, out_sp = O.strided_rulebook(O.input_layer_rules(coords)[1], (2, 2, 2), (2, 2, 2), grid)
    assert tuple(int(v) for v in p.spatial_size) == tuple(out_sp)
    loc = p.get_spatial_locations().numpy()
    og, orr = np.lexsort(loc.T[::-1]), np.lexsort(out_coords.T[::-1])       # GPU: first-appearance rows; oracle: by key
    assert np.array_equal(loc[og], out_coords[orr])
    counts = np.zeros(out_coords.shape[0])
    for r in rules:
        if len(r):
            np.add.at(counts, np.asarray(r)[:, 1], 1)
    assert np.array_equal(p.features[:, 0].float().cpu().numpy()[og], (counts / 8).astype(np.float32)[orr])   # exact in fp32


def test_average_pooling_empty_input(scn):
    e = scn.InputLayer(3, 16)((torch.zeros(0, 4).long().cuda(), torch.zeros(0, 2).cuda(), 2))
    assert scn.AveragePooling(3, 2, 2)(e).features.shape == (0, 2)
