"""Generates tests/golden/legacy*.npz: the reference's LEGACY network files
(src/networks/torch/sparseresnet3d.py and src/networks/torch/sparseresnet.py) imported verbatim
from /root/reference (read-only) on the oracle ``sparseconvnet`` shim, with a reconstructed ``args.network`` config
(the config module they were written against no longer exists in the reference tree).  Build container only:

    python tests/golden/make_golden_legacy.py

Also asserts that sparseeventid_b200/legacy_networks.py is the same composition (same state_dict keys, bit-identical
outputs and gradients on the oracle).  (PARITY UNPINNED vs SCN itself, as for every fixture here.)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import init_deterministic, legacy_batch, LEGACY_CASES  # noqa: E402


def load_reference_module(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle import sparseconvnet_oracle as oscn
    from sparseeventid_b200 import legacy_networks as mirror
    from sparseeventid_b200 import networks, synthetic
    sys.modules["sparseconvnet"] = oscn
    ref3d = load_reference_module("src/networks/torch/sparseresnet3d.py", "ref_sparseresnet3d")
    ref2d = load_reference_module("src/networks/torch/sparseresnet.py", "ref_sparseresnet")
    for case, spec in LEGACY_CASES.items():
        cfg = mirror.LegacyNetworkConfig(**spec["cfg"])
        args = types.SimpleNamespace(network=types.SimpleNamespace(**spec["cfg"]))
        ref_cls = ref3d.ResNet if spec["kind"] == "3d" else ref2d.ResNet
        mir_cls = mirror.LegacyResNet3D if spec["kind"] == "3d" else mirror.LegacyResNet2D
        batch = legacy_batch(case)
        labels = {k: torch.as_tensor(v) for k, v in synthetic.make_labels(2, seed=11).items()}
        results = []
        for build in (lambda: ref_cls(mirror.LEGACY_OUTPUT_SHAPE, args),
                      lambda: mir_cls(oscn, mirror.LEGACY_OUTPUT_SHAPE, cfg)):
            torch.manual_seed(0)
            model = build()
            init_deterministic(model)
            model.train()
            logits = model((torch.as_tensor(batch[0]), torch.as_tensor(batch[1]), batch[2]))
            loss = networks.focal_loss(labels, logits)
            loss.backward()
            results.append((model, logits, loss))
        (m0, l0, s0), (m1, l1, s1) = results
        assert list(m0.state_dict().keys()) == list(m1.state_dict().keys()), "state_dict keys differ"
        assert float(s0.detach()) == float(s1.detach())
        for k in l0:
            assert torch.equal(l0[k], l1[k]), k
        for (n0, p0), (n1, p1) in zip(m0.named_parameters(), m1.named_parameters()):
            assert n0 == n1 and torch.equal(p0.grad, p1.grad), n0
        out = {"n_voxels": np.asarray([batch[0].shape[0]]), "loss": np.asarray([float(s0.detach())])}
        for k, v in l0.items():
            out["logits_" + k] = v.detach().double().numpy()
        out["param_names"] = np.asarray([n for n, _ in m0.named_parameters()])
        out["grad_norms"] = np.asarray([float(p.grad.double().norm()) for _, p in m0.named_parameters()])
        out["n_params"] = np.asarray([sum(p.numel() for p in m0.parameters())])
        path = os.path.join(HERE, f"{case}.npz")
        np.savez_compressed(path, **out)
        print(case, "voxels", out["n_voxels"], "loss", out["loss"], "params", out["n_params"], "->", path,
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
