"""Host-side mirror of the reference's sparse ResNet encoder and classification heads.

The reference model files (``src/networks/resnet.py``, ``sparse_building_blocks.py``,
``classification_head.py``) run unmodified on top of the drop-in ``sparseconvnet`` package, but they
import hydra-registered config enums (``src/config/network.py``) and live outside this repo, so the
GPU box cannot import them.  This module restates the same composition -- same module tree, hence the
same ``state_dict`` keys, same hyper-parameter defaults (src/config/network.py:23-38)
-- for benchmarks, smoke runs and parity tests.  ``scn`` is injected so the identical definition can be
instantiated on the product package or (tests only) on the oracle shim.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence

import torch
from torch import nn


@dataclass
class EncoderConfig:
    """src/config/network.py:23-38 (Repr + ConvRepresentation defaults)."""
    depth: int = 5
    n_initial_filters: int = 32
    n_output_filters: int = 128
    batch_norm: bool = True
    bias: bool = True
    blocks_per_layer: int = 4
    residual: bool = True
    filter_size: int = 3
    multiplicative_growth: bool = False


class Block(nn.Module):
    """conv -> [BatchNormalization] -> activation  (sparse_building_blocks.py:18-57)."""

    def __init__(self, scn, *, nIn, nOut, dim, cfg, activation=None):
        super().__init__()
        kernel = [1, cfg.filter_size, cfg.filter_size] if dim == 2 else [cfg.filter_size] * 3
        self.conv1 = scn.SubmanifoldConvolution(dimension=3, nIn=nIn, nOut=nOut, filter_size=kernel, bias=cfg.bias)
        self._do_normalization = cfg.batch_norm
        if cfg.batch_norm:
            self.norm = scn.BatchNormalization(nOut)
        self.activation = (activation or scn.LeakyReLU)()

    def forward(self, x):
        out = self.conv1(x)
        if self._do_normalization:
            out = self.norm(out)
        return self.activation(out)


class ResidualBlock(nn.Module):
    """Block, Block(Identity), AddTable, LeakyReLU  (sparse_building_blocks.py:61-100)."""

    def __init__(self, scn, *, nIn, nOut, dim, cfg):
        super().__init__()
        self.convolution_1 = Block(scn, nIn=nIn, nOut=nOut, dim=dim, cfg=cfg)
        self.convolution_2 = Block(scn, nIn=nIn, nOut=nOut, dim=dim, cfg=cfg, activation=scn.Identity)
        self.residual = scn.Identity()
        self.relu = scn.LeakyReLU()
        self.add = scn.AddTable()

    def forward(self, x):
        residual = self.residual(x)
        out = self.convolution_2(self.convolution_1(x))
        return self.relu(self.add([out, residual]))


class ConvolutionDownsample(nn.Module):
    """Convolution f=s=2 (no bias) -> BatchNormalization -> LeakyReLU  (sparse_building_blocks.py:103-139)."""

    def __init__(self, scn, *, nIn, nOut, dim, cfg):
        super().__init__()
        f = [1, 2, 2] if dim == 2 else [2, 2, 2]
        self.conv = scn.Convolution(dimension=3, nIn=nIn, nOut=nOut, filter_size=f, filter_stride=f, bias=False)
        self._do_normalization = cfg.batch_norm
        if cfg.batch_norm:
            self.norm = scn.BatchNormalization(nOut)
        self.relu = scn.LeakyReLU()

    def forward(self, x):
        out = self.conv(x)
        if self._do_normalization:
            out = self.norm(out)
        return self.relu(out)


class BlockSeries(nn.Module):
    """n_blocks (Residual)Blocks registered as block_{i}  (sparse_building_blocks.py:231-264)."""

    def __init__(self, scn, *, nIn, n_blocks, dim, cfg):
        super().__init__()
        kind = ResidualBlock if cfg.residual else Block
        self.blocks = [kind(scn, nIn=nIn, nOut=nIn, dim=dim, cfg=cfg) for _ in range(n_blocks)]
        for i, b in enumerate(self.blocks):
            self.add_module(f"block_{i}", b)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class Encoder(nn.Module):
    """src/networks/resnet.py:10-161, sparse mode.  image_size = (1,1024,512,1280) for dune3d
    (dimension 3) or (3,1536,1024) for dune2d (dimension 2), as larcv_dataset.image_size() returns."""

    def __init__(self, scn, cfg: EncoderConfig, image_size: Sequence[int], dimension: int):
        super().__init__()
        nf = cfg.n_initial_filters
        if dimension == 2:
            self.input_layer = scn.InputLayer(dimension=3, spatial_size=torch.tensor(list(image_size)))
            stem = [1, 5, 5]
        else:
            self.input_layer = scn.InputLayer(dimension=3, spatial_size=torch.tensor(list(image_size[1:])))
            stem = [5, 5, 5]
        self.initial_convolution = scn.SubmanifoldConvolution(dimension=3, nIn=1, nOut=nf, filter_size=stem,
                                                              bias=cfg.bias)
        self.network_layers = nn.ModuleList()
        cur = nf
        for _ in range(cfg.depth):
            self.network_layers.append(BlockSeries(scn, nIn=cur, n_blocks=cfg.blocks_per_layer, dim=dimension, cfg=cfg))
            nxt = cur * 2 if cfg.multiplicative_growth else cur + nf
            self.network_layers.append(ConvolutionDownsample(scn, nIn=cur, nOut=nxt, dim=dimension, cfg=cfg))
            cur = nxt
        self.final_layer = BlockSeries(scn, nIn=cur, n_blocks=cfg.blocks_per_layer, dim=dimension, cfg=cfg)
        self.bottleneck = scn.SubmanifoldConvolution(dimension=3, nIn=cur, nOut=cfg.n_output_filters, filter_size=1,
                                                     bias=cfg.bias)
        final_shape = [i // 2 ** cfg.depth for i in image_size]
        if dimension == 2:
            final_shape[0] = 3
        else:
            final_shape = final_shape[1:]
        self.output_shape = [cfg.n_output_filters] + final_shape
        self.pool = nn.Sequential(scn.SparseToDense(dimension=3, nPlanes=self.output_shape[0]))

    def forward(self, x):
        x = self.input_layer(x)
        x = self.initial_convolution(x)
        for layer in self.network_layers:
            x = layer(x)
        x = self.final_layer(x)
        x = self.bottleneck(x)
        x = self.pool(x)
        return torch.tanh(x)


def create_final_dense_chain(spatial_shape, n_in, n_out):
    """classification_head.py:19-28."""
    return nn.Sequential(
        nn.AvgPool3d(spatial_shape),
        nn.Flatten(start_dim=1, end_dim=-1),
        nn.Linear(in_features=n_in, out_features=256),
        nn.Dropout(),
        nn.LeakyReLU(),
        nn.Linear(in_features=256, out_features=n_out),
    )


class MultiHeadOutput(nn.Module):
    """classification_head.py:6-17 (multi_head_output)."""

    def __init__(self, spatial_shape, n_in, output_shape: Dict[str, int]):
        super().__init__()
        self.classification_head = nn.ModuleDict(
            {k: create_final_dense_chain(spatial_shape, n_in, v) for k, v in output_shape.items()})

    def forward(self, x):
        return {k: head(x) for k, head in self.classification_head.items()}


OUTPUT_SHAPE = {"labelneutID": 3, "labelprotID": 3, "labelnpiID": 2, "labelcpiID": 2}   # supervised_eventID.py:224-229
IMAGE_SIZE = {"dune3d": (1, 1024, 512, 1280), "dune2d": (3, 1536, 1024)}               # larcv_fetcher.py:23-48,384-386
DIMENSION = {"dune3d": 3, "dune2d": 2}


def build_networks(scn, dataset: str = "dune3d", cfg: EncoderConfig | None = None, image_size=None):
    """classification_head.py:30-55: (encoder, heads) for a recipe (recipes/dune2d.yaml, dune3d.yaml)."""
    cfg = cfg or EncoderConfig()
    image_size = image_size or IMAGE_SIZE[dataset]
    enc = Encoder(scn, cfg, image_size, DIMENSION[dataset])
    head = MultiHeadOutput(enc.output_shape[1:], enc.output_shape[0], OUTPUT_SHAPE)
    return enc, head


def focal_loss(all_labels, all_logits):
    """supervised_eventID.py:168-196, focal branch (the default, src/config/optimizer.py:49)."""
    loss = 0.0
    for key, logits in all_logits.items():
        y = torch.nn.functional.one_hot(all_labels[key], logits.size(-1))
        p = torch.nn.functional.softmax(logits, dim=-1).clamp(1e-7, 1.0 - 1e-7)
        term = -y * torch.log(p) * (1 - p) ** 2
        loss = loss + term.sum(dim=-1).mean()
    return loss


class EventIDModel(nn.Module):
    """encoder + heads, as supervised_eventID.forward composes them (supervised_eventID.py:53-59)."""

    def __init__(self, encoder, head):
        super().__init__()
        self.encoder = encoder
        self.head = head

    def forward(self, batch):
        return self.head(self.encoder(batch))
