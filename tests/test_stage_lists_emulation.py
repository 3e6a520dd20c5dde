"""CPU check of the LOGIC of the experimental stage-list convolution (csrc/stage_lists.cuh, csrc/conv_tcl.cu), which has
not run on a GPU yet: a numpy emulation that follows the builder and the kernel step by step -- list layout and
padding, stage order (centre step first), PAIR halves, 64-channel chunks, disable-lane masks, which MMA initialises the
accumulator -- with every byte the kernel never writes modelled as NaN (stale shared memory, stale TMEM).  If the
emulated output equals the oracle convolution and holds no NaN, the algorithm is right; what remains for the GPU is
the CUDA-level plumbing (barriers, descriptors, copies), which the kernel shares with the validated k_conv_tc.
"""
import numpy as np
import pytest
import torch

from helpers import blob_sites
from oracle import scn_oracle as O

TILE = 128


def neighbour_table(coords, filt):
    """nbr[k][o] = input row reaching output row o through offset k (-1: none), padded to 128 rows"""
    n = coords.shape[0]
    rules = O.submanifold_rulebook(coords, filt)
    n_pad = (n + TILE - 1) // TILE * TILE
    nbr = -np.ones((len(rules), n_pad), np.int64)
    for k, r in enumerate(rules):
        r = np.asarray(r).reshape(-1, 2)
        nbr[k, r[:, 1]] = r[:, 0]
    return nbr, rules


def build_lists(nbr):
    """k_stage_lists: per (tile, k) the live rows as (source row, row in tile), padded to 8 by repeating the last
    entry, and the mask (True = output row has no neighbour)"""
    K, n_pad = nbr.shape
    lists = {}
    for tile in range(n_pad // TILE):
        for k in range(K):
            j = nbr[k, tile * TILE:(tile + 1) * TILE]
            ent = [(int(j[r]), int(r)) for r in np.nonzero(j >= 0)[0]]
            if ent:
                ent += [ent[-1]] * (-len(ent) % 8)
            lists[(tile, k)] = (ent, j < 0)
    return lists


def emulate(x, W, bias, nbr, n_rows):
    """k_conv_tcl, one tile at a time.  x [n, Cin], W [K, Cin, Cout] (float64; the bf16 rounding is not the point)"""
    K, cin, cout = W.shape
    lists = build_lists(nbr)
    pair = cin == 32
    nch = (cin + 63) // 64
    last_kc = cin - (nch - 1) * 64
    c = (K - 1) // 2
    cstep, chalf = (c >> 1, c & 1) if pair else (c, 0)
    nstep = (K + 1) // 2 if pair else K

    def nat(step):
        return cstep if step == 0 else (step - 1 if step <= cstep else step)

    out = np.full((n_rows, cout), np.nan)
    all_dead = np.ones(TILE, bool)
    for tile in range(nbr.shape[1] // TILE):
        acc = np.full((TILE, cout), np.nan)                      # TMEM is never cleared
        for q in range(nstep * nch):
            step, cc = (q, 0) if pair else (q // nch, q % nch)
            ns = nat(step)
            A = np.full((TILE, 64), np.nan)                      # the A slot holds whatever the previous stage left
            if pair:
                k0, k1 = 2 * ns, 2 * ns + 1
                halves = [(k0, 0)] + ([(k1, 32)] if k1 < K else [])
                for k, col in halves:
                    for src, r in lists[(tile, k)][0]:
                        A[r, col:col + 32] = x[src]
                B = np.full((64, cout), np.nan)                  # weight image tile of pair ns: W[k0] over W[k1]
                B[:32] = W[k0]
                if k1 < K:
                    B[32:] = W[k1]
                masks = [lists[(tile, k0)][1], lists[(tile, k1)][1] if k1 < K else all_dead]
                nk = 4 if k1 < K else 2
                first = q == 0
                if first:                                        # the centre half is unmasked
                    masks[chalf] = np.zeros(TILE, bool)
                for kk in range(nk):
                    kx = kk ^ 2 if (first and chalf) else kk
                    live = ~masks[kx >> 1]
                    sl = slice(16 * kx, 16 * kx + 16)
                    contrib = A[live][:, sl] @ B[sl]
                    acc[live] = contrib if (q == 0 and kk == 0) else acc[live] + contrib
            else:
                k = ns
                width = last_kc if cc == nch - 1 else 64
                for src, r in lists[(tile, k)][0]:
                    A[r, :width] = x[src, cc * 64:cc * 64 + width]
                B = W[k, cc * 64:cc * 64 + width]                # image tile k * nch + cc
                mask = np.zeros(TILE, bool) if q == 0 else lists[(tile, k)][1]
                for kk in range(width // 16):
                    live = ~mask
                    sl = slice(16 * kk, 16 * kk + 16)
                    contrib = A[live][:, sl] @ B[sl]
                    acc[live] = contrib if (q == 0 and kk == 0) else acc[live] + contrib
        lo = tile * TILE
        hi = min(lo + TILE, n_rows)
        out[lo:hi] = acc[:hi - lo] + (0 if bias is None else bias)
    return out


@pytest.mark.parametrize("filt", [(3, 3, 3), (1, 3, 3), (1, 5, 5)])
@pytest.mark.parametrize("cin,cout", [(32, 32), (32, 64), (64, 32), (96, 96), (160, 32), (192, 64)])
def test_stage_list_algorithm_equals_oracle(cin, cout, filt):
    grid, B = (14, 12, 16), 2
    coords = blob_sites(330, grid, B, seed=41)                   # 3 tiles, the last one partial
    n = coords.shape[0]
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n, cin))
    K = int(np.prod(filt))
    W = rng.standard_normal((K, cin, cout))
    bias = rng.standard_normal(cout)
    nbr, rules = neighbour_table(coords, filt)
    assert np.array_equal(nbr[(K - 1) // 2, :n], np.arange(n))   # the centre offset is the identity: what the kernel relies on
    got = emulate(x, W, bias, nbr, n)
    want = O.conv_forward(torch.from_numpy(x), torch.from_numpy(W), torch.from_numpy(bias), rules, n).numpy()
    assert not np.isnan(got).any(), "a stale (never written) value reached an output row"
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)


def test_dgrad_through_the_same_lists():
    """dgrad = the same gather with B_k = W[K-1-k]^T through the SAME table (mirror symmetry), hence the same lists"""
    grid, B, cin, cout, filt = (12, 12, 12), 2, 32, 64, (3, 3, 3)
    coords = blob_sites(200, grid, B, seed=42)
    n = coords.shape[0]
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.standard_normal((n, cin)))
    W = torch.from_numpy(rng.standard_normal((27, cin, cout)))
    dout = rng.standard_normal((n, cout))
    nbr, rules = neighbour_table(coords, filt)
    dx_want, _, _ = O.conv_backward(x, W, False, rules, torch.from_numpy(dout))
    Wt = np.stack([W[26 - k].numpy().T for k in range(27)])      # image of scn_conv_prep_weights(transpose=1, mirror=1)
    dx = emulate(dout, Wt, None, nbr, n)
    assert np.allclose(dx, dx_want.numpy(), rtol=1e-9, atol=1e-9)
