"""Pins the CPU oracle (oracle/scn_oracle.py) WITHOUT SparseConvNet:

  * submanifold conv  == conv3d(densified, padding=f//2) sampled at the active sites
  * strided f==s conv == conv3d(densified, stride=s) everywhere; active outputs == any-in-window
  * BatchNormalization == F.batch_norm (momentum 1-0.9, eps 1e-4) + leaky
  * explicit backward formulas == torch autograd of the dense formulation (float64)
  * vectorised rulebooks == brute-force dictionary rulebooks
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import scn_oracle as O
from helpers import blob_sites, random_sites

torch.manual_seed(0)


def densify(x, coords, grid, batch):
    return O.sparse_to_dense_forward(x, coords, grid, batch)


def torch_weight(w, f):
    # W[k(a,b,c), i, o] -> Wt[o, i, a, b, c]
    K, cin, cout = w.shape
    return w.view(f[0], f[1], f[2], cin, cout).permute(4, 3, 0, 1, 2).contiguous()


@pytest.mark.parametrize("filt", [(3, 3, 3), (1, 3, 3), (5, 5, 5), (1, 5, 5), (1, 1, 1), (3, 1, 5)])
def test_submanifold_rulebook_vs_bruteforce(filt):
    coords = blob_sites(150, (12, 10, 14), 2, seed=3)
    rules = O.submanifold_rulebook(coords, filt)
    brute = O.submanifold_rulebook_bruteforce(coords, filt)
    assert len(rules) == int(np.prod(filt))
    for r, b in zip(rules, brute):
        assert sorted(map(tuple, r.tolist())) == sorted(b)
    # mirror symmetry + identity centre (App. A.3)
    K = len(rules)
    for k in range(K):
        a = set(map(tuple, rules[k].tolist()))
        m = set((o, i) for i, o in rules[K - 1 - k].tolist())
        assert a == m
    centre = rules[(K - 1) // 2]
    assert np.array_equal(centre[:, 0], centre[:, 1]) and centre.shape[0] == coords.shape[0]


@pytest.mark.parametrize("filt", [(3, 3, 3), (1, 3, 3), (5, 5, 5)])
def test_submanifold_conv_dense_identity(filt):
    grid, B, cin, cout = (9, 8, 10), 2, 3, 4
    coords = blob_sites(120, grid, B, seed=5)
    n = coords.shape[0]
    x = torch.randn(n, cin, dtype=torch.float64)
    K = int(np.prod(filt))
    w = torch.randn(K, cin, cout, dtype=torch.float64)
    bias = torch.randn(cout, dtype=torch.float64)
    rules = O.submanifold_rulebook(coords, filt)
    out = O.conv_forward(x, w, bias, rules, n)
    X = densify(x, coords, grid, B)
    Y = F.conv3d(X, torch_weight(w, filt), bias, padding=tuple(f // 2 for f in filt))
    ref = O.sparse_to_dense_backward(Y, coords)
    assert torch.allclose(out, ref, atol=1e-10)


@pytest.mark.parametrize("filt", [(2, 2, 2), (1, 2, 2)])
def test_strided_conv_dense_identity(filt):
    grid, B, cin, cout = (8, 6, 10), 2, 3, 5
    coords = random_sites(90, grid, B, seed=7)
    n = coords.shape[0]
    x = torch.randn(n, cin, dtype=torch.float64)
    K = int(np.prod(filt))
    w = torch.randn(K, cin, cout, dtype=torch.float64)
    out_coords, rules, out_sp = O.strided_rulebook(coords, filt, filt, grid)
    assert sum(len(r) for r in rules) == n          # f == s: each input maps to one output
    out = O.conv_forward(x, w, None, rules, out_coords.shape[0])
    X = densify(x, coords, grid, B)
    Y = F.conv3d(X, torch_weight(w, filt), None, stride=filt)
    dense_out = O.sparse_to_dense_forward(out, out_coords, out_sp, B)
    assert tuple(Y.shape[2:]) == out_sp
    assert torch.allclose(dense_out, Y, atol=1e-10)
    # active output set == any active input in the window
    A = densify(torch.ones(n, 1, dtype=torch.float64), coords, grid, B)
    occ = F.max_pool3d(A, kernel_size=filt, stride=filt)[:, 0] > 0
    got = torch.zeros_like(occ)
    oc = torch.as_tensor(out_coords)
    got[oc[:, 3], oc[:, 0], oc[:, 1], oc[:, 2]] = True
    assert torch.equal(occ, got)
    # sorted-by-key output order (this build's deterministic choice, App. A.4)
    k = O.pack_keys(out_coords)
    assert np.all(np.diff(k) > 0)


def test_strided_general_filter_gt_stride():
    grid, B = (9, 7, 11), 1
    filt, stride = (3, 3, 3), (2, 2, 2)
    coords = random_sites(60, grid, B, seed=9)
    n = coords.shape[0]
    x = torch.randn(n, 2, dtype=torch.float64)
    w = torch.randn(27, 2, 3, dtype=torch.float64)
    out_coords, rules, out_sp = O.strided_rulebook(coords, filt, stride, grid)
    out = O.conv_forward(x, w, None, rules, out_coords.shape[0])
    Y = F.conv3d(densify(x, coords, grid, B), torch_weight(w, filt), None, stride=stride)
    assert torch.allclose(O.sparse_to_dense_forward(out, out_coords, out_sp, B), Y, atol=1e-10)


def test_conv_backward_matches_autograd():
    grid, B, cin, cout, filt = (7, 7, 7), 2, 3, 4, (3, 3, 3)
    coords = blob_sites(80, grid, B, seed=11)
    n = coords.shape[0]
    x = torch.randn(n, cin, dtype=torch.float64, requires_grad=True)
    w = torch.randn(27, cin, cout, dtype=torch.float64, requires_grad=True)
    b = torch.randn(cout, dtype=torch.float64, requires_grad=True)
    rules = O.submanifold_rulebook(coords, filt)
    X = densify(x, coords, grid, B)
    Y = F.conv3d(X, torch_weight(w, filt), b, padding=1)
    y = O.sparse_to_dense_backward(Y, coords)
    dout = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), dout)
    dx, dw, db = O.conv_backward(x.detach(), w.detach(), True, rules, dout)
    assert torch.allclose(dx, gx, atol=1e-10)
    assert torch.allclose(dw, gw, atol=1e-10)
    assert torch.allclose(db, gb, atol=1e-10)


def test_deconvolution_is_transpose_of_convolution():
    grid, B, filt = (8, 8, 8), 1, (2, 2, 2)
    coords = random_sites(70, grid, B, seed=13)
    out_coords, rules, _ = O.strided_rulebook(coords, filt, filt, grid)
    n, m = coords.shape[0], out_coords.shape[0]
    w = torch.randn(8, 3, 3, dtype=torch.float64)
    xf = torch.randn(n, 3, dtype=torch.float64)
    yc = torch.randn(m, 3, dtype=torch.float64)
    conv = O.conv_forward(xf, w, None, rules, m)
    wt = w.transpose(1, 2).contiguous()
    deconv = O.conv_forward(yc, wt, None, O.swap_rules(rules), n)
    # <conv(x), y> == <x, deconv_{W^T}(y)>
    assert torch.allclose((conv * yc).sum(), (xf * deconv).sum(), atol=1e-9)


@pytest.mark.parametrize("leak", [1.0, 0.0, 0.333, 1.0 / 3.0])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_matches_torch(leak, training):
    n, c = 57, 6
    x = torch.randn(n, c, dtype=torch.float64) * 2 + 0.5
    g = torch.rand(c, dtype=torch.float64) + 0.5
    b = torch.randn(c, dtype=torch.float64)
    rm, rv = torch.randn(c, dtype=torch.float64), torch.rand(c, dtype=torch.float64) + 0.5
    rm2, rv2 = rm.clone(), rv.clone()
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.batch_norm(xr, rm2, rv2, gr, br, training, momentum=1 - 0.9, eps=1e-4)
    ref = F.leaky_relu(ref, leak) if leak != 1.0 else ref
    out, mean, invstd = O.batchnorm_forward(x, g, b, rm, rv, training, leakiness=leak)
    assert torch.allclose(out, ref, atol=1e-12)
    assert torch.allclose(rm, rm2, atol=1e-12) and torch.allclose(rv, rv2, atol=1e-12)
    dout = torch.randn_like(out)
    gx, gg, gb = torch.autograd.grad(ref, (xr, gr, br), dout)
    dx, dg, db = O.batchnorm_backward(x, out, g, mean, invstd, dout, training, leak)
    assert torch.allclose(dx, gx, atol=1e-10)
    assert torch.allclose(dg, gg, atol=1e-10)
    assert torch.allclose(db, gb, atol=1e-10)


def test_input_layer_first_appearance_and_sum():
    coords = np.array([[5, 1, 1, 0], [2, 2, 2, 1], [5, 1, 1, 0], [0, 0, 0, 0], [2, 2, 2, 1], [5, 1, 1, 1]])
    feats = torch.arange(6, dtype=torch.float64)[:, None] + 1
    rows, active = O.input_layer_rules(coords)
    assert rows.tolist() == [0, 1, 0, 2, 1, 3]
    assert active.tolist() == [[5, 1, 1, 0], [2, 2, 2, 1], [0, 0, 0, 0], [5, 1, 1, 1]]
    out = O.input_layer_forward(feats, rows, 4, mode=3)
    assert out[:, 0].tolist() == [4.0, 7.0, 4.0, 6.0]
    out4 = O.input_layer_forward(feats, rows, 4, mode=4)
    assert out4[:, 0].tolist() == [2.0, 3.5, 4.0, 6.0]
    assert O.input_layer_forward(feats, rows, 4, mode=1)[:, 0].tolist() == [3.0, 5.0, 4.0, 6.0]
    assert O.input_layer_forward(feats, rows, 4, mode=2)[:, 0].tolist() == [1.0, 2.0, 4.0, 6.0]
    back = O.input_layer_backward(out, rows)
    assert back[:, 0].tolist() == [4.0, 7.0, 4.0, 4.0, 7.0, 6.0]
    assert torch.equal(O.output_layer_forward(out, rows), back)


def test_sparse_to_dense_roundtrip_and_empty():
    grid, B = (4, 5, 6), 3
    coords = random_sites(40, grid, B, seed=17)
    x = torch.randn(coords.shape[0], 7, dtype=torch.float64)
    d = O.sparse_to_dense_forward(x, coords, grid, B)
    assert d.shape == (B, 7, 4, 5, 6)
    assert torch.equal(O.sparse_to_dense_backward(d, coords), x)
    assert float(d.abs().sum()) == pytest.approx(float(x.abs().sum()))
    e = O.sparse_to_dense_forward(torch.zeros(0, 7), np.zeros((0, 4), np.int64), grid, 2)
    assert e.shape == (2, 7, 4, 5, 6) and float(e.abs().sum()) == 0.0


def test_empty_and_single_site_rulebooks():
    r = O.submanifold_rulebook(np.zeros((0, 4), np.int64), (3, 3, 3))
    assert len(r) == 27 and all(len(x) == 0 for x in r)
    one = np.array([[0, 0, 0, 0]])
    r = O.submanifold_rulebook(one, (3, 3, 3))
    assert [len(x) for x in r] == [0] * 13 + [1] + [0] * 13
    # sites on the 16-bit boundary never wrap into a neighbouring field of the packed key
    edge = np.array([[0, 0, 65535, 0], [0, 1, 0, 0], [65535, 65535, 65535, 0], [0, 0, 0, 1]])
    r = O.submanifold_rulebook(edge, (3, 3, 3))
    assert sum(len(x) for x in r) == 4


@pytest.mark.parametrize("pool", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("n_drop", [0, 2])
def test_average_pooling_dense_identity(pool, n_drop):
    """AveragePooling == avg_pool3d of the zero-filled volume (divide by the pool volume), forward and backward."""
    grid, B, c = (8, 6, 10), 2, 5
    coords = random_sites(110, grid, B, seed=11)
    n = coords.shape[0]
    x = torch.randn(n, c, dtype=torch.float64)
    vol = int(np.prod(pool))
    out_coords, rules, out_sp = O.strided_rulebook(coords, pool, pool, grid)
    out = O.average_pooling_forward(x, rules, out_coords.shape[0], vol, n_drop)
    X = densify(x, coords, grid, B).requires_grad_(True)
    Y = F.avg_pool3d(X[:, n_drop:], kernel_size=pool, stride=pool)
    assert torch.allclose(O.sparse_to_dense_forward(out, out_coords, out_sp, B), Y, atol=1e-12)
    dout = torch.randn_like(out)
    dx = O.average_pooling_backward(dout, rules, n, vol, n_drop)
    Y.backward(O.sparse_to_dense_forward(dout, out_coords, out_sp, B))
    assert torch.allclose(dx, O.sparse_to_dense_backward(X.grad, coords), atol=1e-12)
    assert n_drop == 0 or float(dx[:, :n_drop].abs().max()) == 0.0


def test_average_pooling_module_autograd():
    """oracle scn.AveragePooling module: output grid = Convolution's, gradient == explicit formula."""
    import oracle.sparseconvnet_oracle as scn
    grid, B, c = (8, 8, 8), 2, 4
    coords = random_sites(70, grid, B, seed=13)
    feats = torch.randn(coords.shape[0], c, dtype=torch.float64, requires_grad=True)
    t = scn.InputLayer(3, torch.LongTensor(list(grid)), mode=3)((torch.as_tensor(coords), feats, B))
    p = scn.AveragePooling(3, 2, 2)(t)
    assert tuple(int(v) for v in p.spatial_size) == (4, 4, 4)
    q = scn.Convolution(3, c, 3, 2, 2, False).double()(t)
    assert q.features.shape[0] == p.features.shape[0]
    p.features.sum().backward()
    assert torch.allclose(feats.grad, torch.full_like(feats, 1 / 8))
