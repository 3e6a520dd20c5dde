"""Generates tests/golden/*.npz by running the REFERENCE's own model files, imported verbatim from
/root/reference (read-only), on top of the oracle ``sparseconvnet`` shim, with ``hydra`` / ``omegaconf``
stubbed (they only register dataclasses).  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the outputs are committed as small fixtures together
with this script.  What is pinned: the reference's *composition* of the scn layers (Encoder + heads of
recipes/dune3d.yaml and recipes/dune2d.yaml with the default hyper-parameters) evaluated by the oracle
on seeded synthetic events -> logits, loss, encoder output checksums and per-parameter gradient norms.
(The oracle's own arithmetic is pinned separately by tests/test_oracle_dense.py; PARITY UNPINNED vs SCN.)
"""
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
from helpers import init_deterministic, small_batch  # noqa: E402


def install_stubs():
    from oracle import sparseconvnet_oracle as oscn
    sys.modules["sparseconvnet"] = oscn
    hydra = types.ModuleType("hydra")
    core = types.ModuleType("hydra.core")
    cs = types.ModuleType("hydra.core.config_store")

    class ConfigStore:
        _inst = None

        @classmethod
        def instance(cls):
            cls._inst = cls._inst or cls()
            return cls._inst

        def store(self, *a, **k):
            pass

    cs.ConfigStore = ConfigStore
    hydra.core = core
    core.config_store = cs
    sys.modules.update({"hydra": hydra, "hydra.core": core, "hydra.core.config_store": cs})
    om = types.ModuleType("omegaconf")
    om.MISSING = "???"
    sys.modules["omegaconf"] = om
    sys.path.insert(0, REF)


def run_model(encoder, head, batch_tuple, labels, focal_loss):
    coords, feats, bs = batch_tuple
    x = (torch.as_tensor(coords), torch.as_tensor(feats), bs)
    enc = encoder(x)
    logits = head(enc)
    loss = focal_loss(labels, logits)
    loss.backward()
    return enc, logits, loss


def main():
    install_stubs()
    from src.config.framework import DataMode
    from src.config.network import ConvRepresentation
    from src.networks.classification_head import build_networks
    from sparseeventid_b200 import networks as mirror
    from sparseeventid_b200 import synthetic
    from oracle import sparseconvnet_oracle as oscn

    for dataset in ("dune3d", "dune2d"):
        torch.manual_seed(0)
        params = types.SimpleNamespace(
            data=types.SimpleNamespace(dimension=mirror.DIMENSION[dataset]),
            framework=types.SimpleNamespace(mode=DataMode.sparse),
            encoder=ConvRepresentation(),
        )
        image_size = mirror.IMAGE_SIZE[dataset]
        encoder, head = build_networks(params, list(image_size), mirror.OUTPUT_SHAPE)
        model = mirror.EventIDModel(encoder, head)
        init_deterministic(model)
        model.train()
        head.eval()          # Dropout(0.5) off; BatchNormalization stays in training mode
        batch = small_batch(dataset)
        labels = {k: torch.as_tensor(v) for k, v in synthetic.make_labels(2, seed=11).items()}
        enc, logits, loss = run_model(encoder, head, batch, labels, mirror.focal_loss)

        # the mirror in sparseeventid_b200/networks.py must be the same composition: same keys, same numbers
        m_enc, m_head = mirror.build_networks(oscn, dataset)
        m_model = mirror.EventIDModel(m_enc, m_head)
        assert list(m_model.state_dict().keys()) == list(model.state_dict().keys()), "state_dict keys differ"
        init_deterministic(m_model)
        m_model.train()
        m_head.eval()
        enc2, logits2, loss2 = run_model(m_enc, m_head, batch, labels, mirror.focal_loss)
        assert torch.equal(enc, enc2) and float(loss.detach()) == float(loss2.detach()), "mirror != reference composition"
        for (n1, p1), (n2, p2) in zip(model.named_parameters(), m_model.named_parameters()):
            assert n1 == n2 and torch.equal(p1.grad, p2.grad), n1

        enc, loss = enc.detach(), loss.detach()
        out = {
            "n_voxels": np.asarray([batch[0].shape[0]]),
            "loss": np.asarray([float(loss)]),
            "enc_sum": np.asarray([float(enc.double().sum())]),
            "enc_abs_sum": np.asarray([float(enc.double().abs().sum())]),
            "enc_nonzero": np.asarray([int((enc != 0).sum())]),
            "enc_pooled": enc.double().mean(dim=(2, 3, 4)).numpy(),
        }
        for k, v in logits.items():
            out["logits_" + k] = v.detach().double().numpy()
        names, norms, firsts = [], [], []
        for n, p in model.named_parameters():
            names.append(n)
            norms.append(float(p.grad.double().norm()))
            firsts.append(float(p.grad.reshape(-1)[0]))
        out["param_names"] = np.asarray(names)
        out["grad_norms"] = np.asarray(norms)
        out["grad_first"] = np.asarray(firsts)
        out["n_params"] = np.asarray([sum(p.numel() for p in model.parameters())])
        rm = dict(model.named_buffers())
        out["running_mean_l0"] = rm["encoder.network_layers.0.block_0.convolution_1.norm.running_mean"].numpy()
        out["running_var_l0"] = rm["encoder.network_layers.0.block_0.convolution_1.norm.running_var"].numpy()
        path = os.path.join(HERE, f"{dataset}_default_encoder.npz")
        np.savez_compressed(path, **out)
        print(dataset, "voxels", out["n_voxels"], "loss", float(loss), "params", out["n_params"], "->", path,
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
