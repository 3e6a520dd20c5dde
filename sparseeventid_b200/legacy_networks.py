"""Host-side mirror of the reference's legacy sparse ResNets:

  * ``src/networks/torch/sparseresnet3d.py:184-296``  -- full-resolution 3-D network on a 1536^3 grid
    (BASELINE.json configs[4], "preprocess_fullres_3D.cfg shape"),
  * ``src/networks/torch/sparseresnet.py:194-334``    -- 2-D multiplane network on [3, 2048, 1280]: ``depth_pre_merge``
    stages with [1,3,3] kernels (weights shared by the three wire planes, which are stacked along the first
    spatial axis) followed by stages with [3,3,3] kernels that mix the planes.

The reference files run unmodified on the drop-in ``sparseconvnet`` package (tests/golden/make_golden.py imports them
verbatim), but they live outside this repository and read their hyper-parameters from an ``args.network`` config
that no longer exists in the reference tree, so the GPU box needs this restatement: same module tree, hence the
same ``state_dict`` keys; ``scn`` is injected so the identical definition runs on the product package or on the
oracle shim (tests only).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence

import torch
from torch import nn


@dataclass
class LegacyNetworkConfig:
    """``args.network.*`` as read at sparseresnet.py:218-237 / sparseresnet3d.py:209-232.  Defaults follow the DUNE
    production run name ``..._bpl2_nf32_..._bntrue_lrfalse`` (analysis/dune/Network Analysis.ipynb)."""
    n_initial_filters: int = 32
    network_depth: int = 6
    depth_pre_merge: int = 3          # 2-D network only
    res_blocks_per_layer: int = 2
    batch_norm: bool = True
    leaky_relu: bool = False


def _act(scn, leaky):
    return scn.LeakyReLU() if leaky else scn.ReLU()


def _bn_act(scn, planes, leaky):
    return scn.BatchNormLeakyReLU(planes) if leaky else scn.BatchNormReLU(planes)


class SparseResidualBlock(nn.Module):
    """conv1 -> bn1(+act) -> conv2 -> bn2 -> AddTable -> act   (sparseresnet.py:46-105, sparseresnet3d.py:41-99)."""

    def __init__(self, scn, inplanes, outplanes, cfg, kernel):
        super().__init__()
        self.batch_norm = cfg.batch_norm
        self.conv1 = scn.SubmanifoldConvolution(dimension=3, nIn=inplanes, nOut=outplanes, filter_size=kernel, bias=False)
        if cfg.batch_norm:
            self.bn1 = _bn_act(scn, outplanes, cfg.leaky_relu)
        self.conv2 = scn.SubmanifoldConvolution(dimension=3, nIn=outplanes, nOut=outplanes, filter_size=kernel, bias=False)
        if cfg.batch_norm:
            self.bn2 = scn.BatchNormalization(outplanes)
        self.residual = scn.Identity()
        self.relu = _act(scn, cfg.leaky_relu)
        self.add = scn.AddTable()

    def forward(self, x):
        residual = self.residual(x)
        out = self.conv1(x)
        out = self.bn1(out) if self.batch_norm else self.relu(out)
        out = self.conv2(out)
        if self.batch_norm:
            out = self.bn2(out)
        return self.relu(self.add([out, residual]))


class SparseBlock(nn.Module):
    """conv1 -> bn1(+act) | act   (sparseresnet.py:12-43)."""

    def __init__(self, scn, inplanes, outplanes, cfg, kernel):
        super().__init__()
        self.batch_norm = cfg.batch_norm
        self.conv1 = scn.SubmanifoldConvolution(dimension=3, nIn=inplanes, nOut=outplanes, filter_size=kernel, bias=False)
        if cfg.batch_norm:
            self.bn1 = _bn_act(scn, outplanes, cfg.leaky_relu)
        else:
            self.relu = _act(scn, cfg.leaky_relu)

    def forward(self, x):
        out = self.conv1(x)
        return self.bn1(out) if self.batch_norm else self.relu(out)


class SparseConvolutionDownsample(nn.Module):
    """Convolution f == s -> [bn] -> act   (sparseresnet.py:108-136, sparseresnet3d.py:103-131)."""

    def __init__(self, scn, inplanes, outplanes, cfg, size):
        super().__init__()
        self.batch_norm = cfg.batch_norm
        self.conv = scn.Convolution(dimension=3, nIn=inplanes, nOut=outplanes, filter_size=size, filter_stride=size,
                                    bias=False)
        if cfg.batch_norm:
            self.bn = scn.BatchNormalization(outplanes)
        self.relu = _act(scn, cfg.leaky_relu)

    def forward(self, x):
        out = self.conv(x)
        if self.batch_norm:
            out = self.bn(out)
        return self.relu(out)


class SparseBlockSeries(nn.Module):
    """n_blocks blocks registered as block_{i}   (sparseresnet.py:139-159)."""

    def __init__(self, scn, inplanes, n_blocks, cfg, kernel, residual=True):
        super().__init__()
        kind = SparseResidualBlock if residual else SparseBlock
        self.blocks = [kind(scn, inplanes, inplanes, cfg, kernel) for _ in range(n_blocks)]
        for i, b in enumerate(self.blocks):
            self.add_module(f"block_{i}", b)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


def _final(scn, n_filters, cfg, kernel, output_shape):
    return nn.ModuleDict({
        key: nn.Sequential(
            SparseBlockSeries(scn, n_filters, cfg.res_blocks_per_layer, cfg, kernel),
            scn.SubmanifoldConvolution(dimension=3, nIn=n_filters, nOut=output_shape[key][-1], filter_size=1, bias=False),
            scn.SparseToDense(dimension=3, nPlanes=output_shape[key][-1]))
        for key in output_shape})


def _pool_heads(final_layer, x, batch_size):
    """Global average pooling over the dense extent + view   (sparseresnet.py:318-332, sparseresnet3d.py:276-292)."""
    output = {}
    for key in final_layer:
        y = final_layer[key](x)
        kernel_size = y.shape[2:]
        y = torch.squeeze(nn.AvgPool3d(kernel_size, ceil_mode=False)(y))
        output[key] = y.view([batch_size, y.shape[-1]])
    return output


class LegacyResNet3D(nn.Module):
    """sparseresnet3d.ResNet: stem 5^3 1->nf, ``network_depth`` x (residual series, stride-2 downsample nf -> nf+nf0),
    per-key head (series, 1x1, SparseToDense) + global average pool."""

    def __init__(self, scn, output_shape: Dict[str, Sequence[int]], cfg: LegacyNetworkConfig, spatial_size=(1536, 1536, 1536)):
        super().__init__()
        self.input_tensor = scn.InputLayer(dimension=3, spatial_size=tuple(spatial_size))
        nf0 = cfg.n_initial_filters
        self.initial_convolution = scn.SubmanifoldConvolution(3, 1, nf0, filter_size=5, bias=False)
        n_filters = nf0
        self.convolutional_layers = []
        for layer in range(cfg.network_depth):
            self.convolutional_layers.append(SparseBlockSeries(scn, n_filters, cfg.res_blocks_per_layer, cfg, 3))
            out_filters = n_filters + nf0
            self.convolutional_layers.append(SparseConvolutionDownsample(scn, n_filters, out_filters, cfg, 2))
            n_filters = out_filters
            self.add_module(f"conv_{layer}", self.convolutional_layers[-2])
            self.add_module(f"down_{layer}", self.convolutional_layers[-1])
        self.final_layer = _final(scn, n_filters, cfg, 3, output_shape)

    def forward(self, x):
        batch_size = x[2]
        x = self.initial_convolution(self.input_tensor(x))
        for layer in self.convolutional_layers:
            x = layer(x)
        return _pool_heads(self.final_layer, x, batch_size)


class LegacyResNet2D(nn.Module):
    """sparseresnet.ResNet (2-D multiplane): the three planes are the first spatial axis of a [3, H, W] grid;
    pre-merge stages use [1,3,3] kernels (same weights on every plane), post-merge stages [3,3,3] kernels that mix
    the planes; every downsample is [1,2,2] / [1,2,2]."""

    def __init__(self, scn, output_shape: Dict[str, Sequence[int]], cfg: LegacyNetworkConfig, spatial_size=(3, 2048, 1280)):
        super().__init__()
        self.input_tensor = scn.InputLayer(dimension=3, spatial_size=list(spatial_size))
        nf0 = cfg.n_initial_filters
        self.initial_convolution = scn.SubmanifoldConvolution(dimension=3, nIn=1, nOut=nf0, filter_size=[1, 5, 5], bias=False)
        n_filters = nf0
        self.pre_convolutional_layers = nn.ModuleList()
        for _ in range(cfg.depth_pre_merge):
            out_filters = n_filters + nf0
            self.pre_convolutional_layers.append(SparseBlockSeries(scn, n_filters, cfg.res_blocks_per_layer, cfg, [1, 3, 3]))
            self.pre_convolutional_layers.append(SparseConvolutionDownsample(scn, n_filters, out_filters, cfg, [1, 2, 2]))
            n_filters = out_filters
        self.post_convolutional_layers = nn.ModuleList()
        for _ in range(cfg.network_depth - cfg.depth_pre_merge):
            out_filters = n_filters + nf0
            self.post_convolutional_layers.append(SparseBlockSeries(scn, n_filters, cfg.res_blocks_per_layer, cfg, [3, 3, 3]))
            self.post_convolutional_layers.append(SparseConvolutionDownsample(scn, n_filters, out_filters, cfg, [1, 2, 2]))
            n_filters = out_filters
        self.final_layer = _final(scn, n_filters, cfg, [3, 3, 3], output_shape)

    def forward(self, x):
        batch_size = x[-1]
        x = self.initial_convolution(self.input_tensor(x))
        for layer in self.pre_convolutional_layers:
            x = layer(x)
        for layer in self.post_convolutional_layers:
            x = layer(x)
        return _pool_heads(self.final_layer, x, batch_size)


LEGACY_OUTPUT_SHAPE = {"labelneutID": [3], "labelprotID": [3], "labelnpiID": [2], "labelcpiID": [2]}
