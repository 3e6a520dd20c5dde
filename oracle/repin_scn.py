"""Re-pin the oracle against REAL SparseConvNet the moment it is available (test infrastructure).

The oracle (oracle/scn_oracle.py) restates SparseConvNet's published algorithm from recollection: the reference pins no
version, ships no tests and its dependency is absent from this image, so parity is "unpinned" (DESIGN.md 2).  This
script closes that gap without any further work once `import sparseconvnet` resolves to the real package -- e.g. after
    python -m pip install --no-index --no-build-isolation --target baseline/_ref <SparseConvNet checkout>
(baseline/_ref is git-ignored for exactly this purpose):

    python oracle/repin_scn.py            # exit 0: every check within its bar; 1: a convention differs; 2: SCN not found

What it does, all on the CPU through SparseConvNet's own code path:
  1. layer conventions the dense identities cannot arbitrate, one by one, oracle vs SCN on seeded inputs: InputLayer
     row numbering (mode 3, duplicates summed), submanifold / strided rulebooks (order-normalised, bit-exact),
     weight layout [K,1,Cin,Cout] and offset order, BatchNormalization eps / momentum / running-variance convention,
     LeakyReLU default leakiness, SparseToDense layout;
  2. the reference's own model files (src/networks/resnet.py:10-161 via classification_head.py:30-55), imported
     verbatim on real SCN, regenerate the quantities stored in tests/golden/*_default_encoder.npz, which are diffed
     against the committed fixtures at 2e-3 relative.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
TOL = 2e-3


def find_real_scn():
    """The real package has the compiled `sparseconvnet.SCN` extension; the repo's drop-in (ROOT/sparseconvnet) has not."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), None):
        path = list(sys.path)
        try:
            if cand is not None:
                if not os.path.isdir(cand):
                    continue
                sys.path.insert(0, cand)
            else:
                sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
            sys.modules.pop("sparseconvnet", None)
            mod = importlib.import_module("sparseconvnet")
            if hasattr(mod, "SCN") and "sparseeventid_b200" not in getattr(mod.SubmanifoldConvolution, "__module__", ""):
                return mod
        except Exception:
            pass
        finally:
            sys.path = path
        sys.modules.pop("sparseconvnet", None)
    return None


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def layer_checks(scn):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import blob_sites
    from oracle import scn_oracle as O
    from oracle import sparseconvnet_oracle as oscn
    oscn.set_numerics("fp32")
    bad = []

    def check(name, ok, detail=""):
        print(("ok   " if ok else "FAIL ") + name + (" " + detail if detail else ""))
        if not ok:
            bad.append(name)

    grid, B = (24, 20, 28), 3
    coords = blob_sites(300, grid, B, seed=5)
    dup = np.concatenate([coords, coords[::7]], 0)                         # duplicates: mode 3 sums them
    feats = torch.randn(dup.shape[0], 2)
    xi = scn.InputLayer(3, torch.LongTensor(list(grid)), mode=3)([torch.as_tensor(dup), feats, B])
    xo = oscn.InputLayer(3, list(grid), mode=3)((torch.as_tensor(dup), feats, B))
    loc_s, loc_o = xi.get_spatial_locations().numpy(), xo.get_spatial_locations().numpy()
    check("InputLayer: first-appearance row numbering", np.array_equal(loc_s, loc_o))
    check("InputLayer: duplicates summed (mode 3)", rel(xi.features.detach(), xo.features.detach()) < 1e-6)

    torch.manual_seed(1)
    for filt in ((3, 3, 3), (1, 3, 3), (5, 5, 5)):
        ms = scn.SubmanifoldConvolution(3, 2, 4, list(filt), True)
        mo = oscn.SubmanifoldConvolution(3, 2, 4, list(filt), True)
        check(f"SubmanifoldConvolution{filt}: weight shape", tuple(ms.weight.shape) == tuple(mo.weight.shape),
              f"{tuple(ms.weight.shape)} vs {tuple(mo.weight.shape)}")
        with torch.no_grad():
            ms.weight.normal_(); ms.bias.normal_()
            mo.weight.copy_(ms.weight.reshape(mo.weight.shape)); mo.bias.copy_(ms.bias)
        ys, yo = ms(xi), mo(xo)
        check(f"SubmanifoldConvolution{filt}: offset order / weight layout / bias", rel(ys.features.detach(), yo.features.detach()) < 1e-5)
    cs, co = scn.Convolution(3, 2, 3, 2, 2, False), oscn.Convolution(3, 2, 3, 2, 2, False)
    with torch.no_grad():
        cs.weight.normal_()
        co.weight.copy_(cs.weight.reshape(co.weight.shape))
    ys, yo = cs(xi), co(xo)
    ls, lo = ys.get_spatial_locations().numpy(), yo.get_spatial_locations().numpy()
    ps, po = np.lexsort(ls.T[::-1]), np.lexsort(lo.T[::-1])
    check("Convolution 2/2: output sites", np.array_equal(ls[ps], lo[po]))
    check("Convolution 2/2: offset order / weight layout", rel(ys.features.detach()[ps], yo.features.detach()[po]) < 1e-5)

    bs, bo = scn.BatchNormalization(2), oscn.BatchNormalization(2)
    check("BatchNormalization: eps / momentum", (bs.eps, bs.momentum) == (bo.eps, bo.momentum), f"{bs.eps},{bs.momentum} vs {bo.eps},{bo.momentum}")
    bs.train(); bo.train()
    ys, yo = bs(xi), bo(xo)
    check("BatchNormalization: training output", rel(ys.features.detach(), yo.features.detach()) < 1e-5)
    check("BatchNormalization: running_mean", rel(bs.running_mean, bo.running_mean) < 1e-5)
    check("BatchNormalization: running_var (unbiased, inverted momentum)", rel(bs.running_var, bo.running_var) < 1e-5)
    ys, yo = scn.LeakyReLU()(xi), oscn.LeakyReLU()(xo)
    check("LeakyReLU: default leakiness", rel(ys.features.detach(), yo.features.detach()) < 1e-6)
    ds, do = scn.SparseToDense(3, 2)(xi), oscn.SparseToDense(3, 2)(xo)
    check("SparseToDense: layout [B, C, *spatial]", tuple(ds.shape) == tuple(do.shape) and rel(ds.detach(), do.detach()) < 1e-6)
    return bad


def network_checks(scn):
    """tests/golden/make_golden.py's computation, with `sparseconvnet` = the real package, diffed against the fixtures."""
    sys.modules["sparseconvnet"] = scn
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as G
    from oracle import sparseconvnet_oracle as oscn
    G.install_stubs()
    sys.modules["sparseconvnet"] = scn                                    # install_stubs points it at the oracle shim
    from helpers import init_deterministic, small_batch
    from src.config.framework import DataMode
    from src.config.network import ConvRepresentation
    from src.networks.classification_head import build_networks
    from sparseeventid_b200 import networks as mirror
    from sparseeventid_b200 import synthetic
    del oscn
    bad = []
    for dataset in ("dune3d", "dune2d"):
        want = np.load(os.path.join(ROOT, "tests", "golden", f"{dataset}_default_encoder.npz"))
        params = types.SimpleNamespace(data=types.SimpleNamespace(dimension=mirror.DIMENSION[dataset]),
                                       framework=types.SimpleNamespace(mode=DataMode.sparse), encoder=ConvRepresentation())
        encoder, head = build_networks(params, list(mirror.IMAGE_SIZE[dataset]), mirror.OUTPUT_SHAPE)
        model = mirror.EventIDModel(encoder, head)
        names = [n for n, _ in model.named_parameters()]
        if names != [str(s) for s in want["param_names"]]:
            print(f"FAIL {dataset}: parameter names differ from the fixture")
            bad.append(dataset + " names")
            continue
        init_deterministic(model)
        model.train(); head.eval()
        labels = {k: torch.as_tensor(v) for k, v in synthetic.make_labels(2, seed=11).items()}
        enc, logits, loss = G.run_model(encoder, head, small_batch(dataset), labels, mirror.focal_loss)
        errs = {"loss": rel(float(loss.detach()), want["loss"]),
                "enc_pooled": rel(enc.detach().double().mean(dim=(2, 3, 4)).numpy(), want["enc_pooled"])}
        for k, v in logits.items():
            errs["logits_" + k] = rel(v.detach().numpy(), want["logits_" + k])
        gn = np.asarray([float(p.grad.double().norm()) for _, p in model.named_parameters()])
        wn = want["grad_norms"]
        big = wn > 1e-3 * wn.max()
        errs["grad_norms"] = float((np.abs(gn - wn)[big] / wn[big]).max())
        ok = all(e <= TOL for e in errs.values())
        print(("ok   " if ok else "FAIL ") + f"{dataset} default encoder + heads on real SparseConvNet vs committed fixture: {errs}")
        if not ok:
            bad.append(dataset)
    return bad


def main():
    scn = find_real_scn()
    if scn is None:
        print("SparseConvNet is not importable (neither baseline/_ref nor site-packages): parity stays UNPINNED.\n"
              "Install it into baseline/_ref (see the header of this file) and run again.")
        return 2
    print("real SparseConvNet:", getattr(scn, "__file__", "?"))
    bad = layer_checks(scn)
    if os.path.isdir(os.path.join(REF, "src")):
        bad += network_checks(scn)
    else:
        print("reference tree absent: network-level re-pin skipped")
    print("PINNED: every convention the oracle restates matches SparseConvNet" if not bad else f"DIFFERENCES: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
