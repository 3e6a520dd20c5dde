"""Checkpoints in the reference's (pytorch_lightning) layout -- SURVEY.md 8f-3, reference src/utils/create_trainer.py:83-115:
``{"state_dict": {"encoder.<...>", "head.<...>"}}``; full restore, and ``restore_encoder_only`` (encoder loaded from
the keys containing "encoder" with the prefix stripped, then frozen).  CPU, oracle modules, a small encoder."""
import torch

from test_dist_trainer import _batch, _make_trainer


def _step(networks, model, opt):
    batch, labels = _batch(0)
    loss = networks.focal_loss(labels, model(batch))
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.detach())


def test_lightning_layout_round_trip(tmp_path):
    from sparseeventid_b200.trainer import checkpoint_dict, restore_checkpoint
    networks, model = _make_trainer(seed=1)
    live = list(model.parameters())
    opt = torch.optim.Adam(live, lr=1e-3, eps=1e-6, betas=(0.8, 0.9))
    model.train()
    _step(networks, model, opt)
    ck = checkpoint_dict(model, opt, None, global_step=1)
    assert set(ck) >= {"state_dict", "global_step", "optimizer_states"}
    keys = list(ck["state_dict"])
    assert any(k.startswith("encoder.network_layers.") for k in keys) and any(k.startswith("head.") for k in keys)
    path = tmp_path / "step=1.ckpt"
    torch.save(ck, path)

    _, other = _make_trainer(seed=2)
    opt2 = torch.optim.Adam(other.parameters(), lr=1e-3, eps=1e-6, betas=(0.8, 0.9))
    assert restore_checkpoint(other, torch.load(path, weights_only=False), optimizer=opt2) == 1
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k
    # the optimizer state came along: one more step from both gives the same parameters
    model.head.eval(); other.head.eval()                  # Dropout off
    la, lb = _step(networks, model, opt), _step(networks, other, opt2)
    assert la == lb
    for a, b in zip(model.parameters(), other.parameters()):
        assert torch.equal(a, b)


def test_restore_encoder_only_freezes_the_encoder():
    from sparseeventid_b200.trainer import checkpoint_dict, restore_checkpoint
    _, trained = _make_trainer(seed=3)
    _, fresh = _make_trainer(seed=4)
    head_before = {k: v.clone() for k, v in fresh.head.state_dict().items()}
    restore_checkpoint(fresh, checkpoint_dict(trained), encoder_only=True)
    for (k, a), (_, b) in zip(trained.encoder.state_dict().items(), fresh.encoder.state_dict().items()):
        assert torch.equal(a, b), k
    for k, v in fresh.head.state_dict().items():          # the heads keep their own initialisation
        assert torch.equal(v, head_before[k]), k
    assert all(not p.requires_grad for p in fresh.encoder.parameters())
    assert all(p.requires_grad for p in fresh.head.parameters())


def test_reference_shaped_checkpoint_loads_and_envelope_is_lightning_compatible():
    """A checkpoint as the reference's LightningModule writes it: extra non-network tensors (criterion.weight with
    loss_balance_scheme=even), the Lightning envelope keys; and this repo's own files carry the same envelope."""
    from sparseeventid_b200.trainer import checkpoint_dict, restore_checkpoint
    import pytest
    _, trained = _make_trainer(seed=5)
    _, fresh = _make_trainer(seed=6)
    ck = checkpoint_dict(trained, global_step=50, epoch=1)
    for k in ("pytorch-lightning_version", "epoch", "global_step", "loops", "callbacks", "state_dict", "optimizer_states",
              "lr_schedulers"):
        assert k in ck, k
    ref_like = dict(ck)
    ref_like["state_dict"] = dict(ck["state_dict"])
    ref_like["state_dict"]["criterion.weight"] = torch.ones(3)              # not part of encoder / head
    ref_like["hyper_parameters"] = {"anything": 1}
    assert restore_checkpoint(fresh, ref_like) == 50
    for (k, a), (_, b) in zip(trained.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k
    broken = dict(ck)
    broken["state_dict"] = {k: v for k, v in ck["state_dict"].items() if "running_mean" not in k}
    with pytest.raises(KeyError):
        restore_checkpoint(fresh, broken)
