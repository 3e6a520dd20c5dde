"""Oracle-backed stand-in for the ``sparseconvnet`` module surface.

TEST INFRASTRUCTURE ONLY (see oracle/scn_oracle.py header; PARITY UNPINNED).
It lets the reference's own model files (``/root/reference/src/networks/*.py``) be
imported verbatim on CPU (``sys.modules['sparseconvnet'] = this``) to generate golden
vectors, and it is the CPU baseline timed by ``bench.py --impl reference``.
The product package never imports it.

Surface restated: every ``scn.*`` symbol the reference touches.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from .. import scn_oracle as O

__all__ = [
    "SparseConvNetTensor", "Metadata", "InputLayer", "OutputLayer", "SubmanifoldConvolution",
    "Convolution", "Deconvolution", "AveragePooling", "BatchNormalization", "BatchNormReLU", "BatchNormLeakyReLU",
    "LeakyReLU", "ReLU", "Tanh", "Sigmoid", "Identity", "AddTable", "SparseToDense", "Sequential",
]


# Stated-numerics emulation (tests only).  The CUDA path has three modes (sparseeventid_b200/scn/config.py):
#   fp32 : nothing rounded                              -> NUMERICS all False
#   mixed: bf16 tensor-core operands, fp32 storage      -> operand_round
#   bf16 : bf16 operands AND bf16 feature storage       -> operand_round + storage_round
# With these flags the oracle does float64 arithmetic on exactly the values the kernels see, so the
# comparison isolates the kernels' own arithmetic from the precision the mode states.
NUMERICS = {"operand_round": False, "storage_round": False, "fuse": True}


def set_numerics(mode: str, fuse: bool = True):
    """fuse: emulate the product's BatchNormalization/AddTable + LeakyReLU fusion, i.e. ONE storage rounding after
    the activation instead of one after each module (only observable when storage_round is on)."""
    NUMERICS["operand_round"] = mode in ("mixed", "bf16")
    NUMERICS["storage_round"] = mode == "bf16"
    NUMERICS["fuse"] = fuse


def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


def _rs(t):
    return _bf16(t) if NUMERICS["storage_round"] else t


def _tensor_core_shape(k, n_in, n_out):
    """Mirror of scn_conv_path() > 0: only these shapes round their operands to bf16."""
    return n_in % 32 == 0 and n_out % 32 == 0 and n_in <= 256 and n_out <= 256 and k <= 128


def _ro(t, k, n_in, n_out):
    if NUMERICS["operand_round"] and _tensor_core_shape(k, n_in, n_out):
        return _bf16(t)
    return t


class _RoundStorage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


def round_storage(t):
    """Differentiable bf16 rounding (identity backward): a torch ``.to(bfloat16)`` cast in the GPU pipeline."""
    return _RoundStorage.apply(t)


class Metadata:
    """Per-forward cache: active sites per spatial size + rulebooks."""

    def __init__(self, dimension=3):
        self.dimension = dimension
        self.levels = {}          # spatial tuple -> int64 [N,4] coords in row order
        self.subm = {}            # (spatial, filter) -> rules
        self.strided = {}         # (in_spatial, filter, stride) -> (out_spatial, rules)
        self.row_of_input = None
        self.batch_size = 0
        self.input_spatial = None

    def get_subm(self, spatial, filt):
        key = (spatial, filt)
        if key not in self.subm:
            self.subm[key] = O.submanifold_rulebook(self.levels[spatial], filt)
        return self.subm[key]

    def get_strided(self, spatial, filt, stride):
        key = (spatial, filt, stride)
        if key not in self.strided:
            out_coords, rules, out_spatial = O.strided_rulebook(self.levels[spatial], filt, stride, spatial)
            if out_spatial not in self.levels:
                self.levels[out_spatial] = out_coords
            else:
                # fine->coarse requested again on an existing coarse grid: re-index to it
                have = self.levels[out_spatial]
                remap = O._lookup(*_sorted(have), O.pack_keys(out_coords))
                rules = [np.stack([r[:, 0], remap[r[:, 1]]], 1) if len(r) else r for r in rules]
            self.strided[key] = (out_spatial, rules)
        return self.strided[key]


def _sorted(coords):
    k = O.pack_keys(coords)
    o = np.argsort(k, kind="stable")
    return k[o], o


class SparseConvNetTensor:
    def __init__(self, features=None, metadata=None, spatial_size=None, pending=None):
        self._features = features
        self._pending = pending       # (run, fuse) of a BatchNormalization / AddTable awaiting a possible LeakyReLU
        self.metadata = metadata
        self.spatial_size = spatial_size

    @property
    def features(self):
        if self._pending is not None:
            self._features = self._pending[0]()
            self._pending = None
        return self._features

    @features.setter
    def features(self, value):
        self._features = value
        self._pending = None

    def get_spatial_locations(self, spatial_size=None):
        sp = tuple(int(v) for v in (self.spatial_size if spatial_size is None else spatial_size))
        return torch.as_tensor(self.metadata.levels[sp], dtype=torch.long)

    def batch_size(self):
        return self.metadata.batch_size

    def cpu(self):
        return self

    def __repr__(self):
        return f"SparseConvNetTensor<oracle>(features={tuple(self.features.shape)}, spatial={self.spatial_size})"


def _sp(t):
    return tuple(int(v) for v in t.spatial_size)


# ---------------------------------------------------------------- autograd glue


class _InputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, rows, n_active, mode):
        ctx.rows, ctx.mode = rows, mode
        return O.input_layer_forward(feats, rows, n_active, mode)

    @staticmethod
    def backward(ctx, dout):
        return O.input_layer_backward(dout.contiguous(), ctx.rows, ctx.mode), None, None, None


class _OutputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, rows):
        ctx.rows, ctx.n = rows, feats.shape[0]
        return O.output_layer_forward(feats, rows)

    @staticmethod
    def backward(ctx, dout):
        return O.output_layer_backward(dout.contiguous(), ctx.rows, ctx.n), None


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, rules, n_out):
        ctx.rules = rules
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight)
        w3 = weight.view(weight.shape[0], weight.shape[-2], weight.shape[-1])
        k, cin, cout = w3.shape
        return _rs(O.conv_forward(_ro(x, k, cin, cout), _ro(w3, k, cin, cout), bias, rules, n_out))

    @staticmethod
    def backward(ctx, dout):
        x, weight = ctx.saved_tensors
        w3 = weight.view(weight.shape[0], weight.shape[-2], weight.shape[-1])
        k, cin, cout = w3.shape
        dx, dw, db = O.conv_backward(_ro(x, k, cin, cout), _ro(w3, k, cin, cout), ctx.has_bias, ctx.rules,
                                     _ro(dout.contiguous(), k, cin, cout))
        if ctx.has_bias:
            db = dout.sum(0)
        return _rs(dx), dw.view_as(weight), db, None, None


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rules, n_out, volume, n_drop):
        ctx.rules, ctx.n_in, ctx.volume, ctx.n_drop = rules, x.shape[0], volume, n_drop
        return _rs(O.average_pooling_forward(x, rules, n_out, volume, n_drop))

    @staticmethod
    def backward(ctx, dout):
        return _rs(O.average_pooling_backward(dout.contiguous(), ctx.rules, ctx.n_in, ctx.volume, ctx.n_drop)), \
            None, None, None, None


class _BNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, rm, rv, training, eps, momentum, leak):
        out, mean, invstd = O.batchnorm_forward(x, gamma, beta, rm, rv, training, eps, momentum, leak)
        ctx.training, ctx.leak = training, leak
        ctx.save_for_backward(x, out, gamma, mean, invstd)
        return _rs(out)

    @staticmethod
    def backward(ctx, dout):
        x, out, gamma, mean, invstd = ctx.saved_tensors
        dx, dg, db = O.batchnorm_backward(x, out, gamma, mean, invstd, dout.contiguous(), ctx.training, ctx.leak)
        return _rs(dx), dg, db, None, None, None, None, None, None


class _LeakyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, leak):
        ctx.leak = leak
        ctx.save_for_backward(x)
        return _rs(O.leaky_relu_forward(x, leak))

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return _rs(O.leaky_relu_backward(x, dout, ctx.leak)), None


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return _rs(a + b)

    @staticmethod
    def backward(ctx, dout):
        return dout, dout


class _AddLeakyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, leak):
        out = _rs(O.leaky_relu_forward(a + b, leak))
        ctx.leak = leak
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        d = _rs(O.leaky_relu_backward(out, dout, ctx.leak))
        return d, d, None


class _S2DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coords, spatial, batch_size):
        ctx.coords = coords
        return O.sparse_to_dense_forward(x, coords, spatial, batch_size)

    @staticmethod
    def backward(ctx, dout):
        return _rs(O.sparse_to_dense_backward(dout, ctx.coords)), None, None, None


# ---------------------------------------------------------------- modules


class InputLayer(nn.Module):
    """scn.InputLayer(dimension, spatial_size, mode=3)  (src/networks/resnet.py:26-29,40-43)."""

    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        self.dimension = dimension
        self.spatial_size = torch.LongTensor(list(O.as_triple(spatial_size, dimension)))
        self.mode = mode

    def forward(self, input):
        coords, feats = input[0], input[1]
        batch_size = int(input[2]) if len(input) > 2 else 0
        coords = torch.as_tensor(coords).cpu().long().numpy()
        if coords.shape[1] == self.dimension:       # single sample: add batch column 0
            coords = np.concatenate([coords, np.zeros((coords.shape[0], 1), np.int64)], 1)
        feats = torch.as_tensor(feats)
        md = Metadata(self.dimension)
        rows, active = O.input_layer_rules(coords, self.mode)
        sp = tuple(int(v) for v in self.spatial_size)
        md.levels[sp] = active
        md.row_of_input = rows
        md.input_spatial = sp
        md.batch_size = max(batch_size, int(coords[:, -1].max()) + 1 if coords.shape[0] else 0)
        out = SparseConvNetTensor(metadata=md, spatial_size=self.spatial_size)
        out.features = _InputFn.apply(feats, rows, active.shape[0], self.mode)
        return out


class OutputLayer(nn.Module):
    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, input):
        return _OutputFn.apply(input.features, input.metadata.row_of_input)


def _init_weight(k, n_in, n_out):
    w = torch.empty(k, 1, n_in, n_out)
    w.normal_(0, math.sqrt(2.0 / (n_in * k)))
    return nn.Parameter(w)


class SubmanifoldConvolution(nn.Module):
    """(dimension, nIn, nOut, filter_size, bias, groups=1)  (sparse_building_blocks.py:29-34)."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        assert groups == 1
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = O.as_triple(filter_size, dimension)
        self.filter_volume = int(np.prod(self.filter_size))
        self.weight = _init_weight(self.filter_volume, nIn, nOut)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        md, sp = input.metadata, _sp(input)
        rules = md.get_subm(sp, self.filter_size)
        out = SparseConvNetTensor(metadata=md, spatial_size=input.spatial_size)
        out.features = _ConvFn.apply(input.features, self.weight, self.bias, rules, input.features.shape[0])
        return out


class Convolution(nn.Module):
    """(dimension, nIn, nOut, filter_size, filter_stride, bias)  (sparse_building_blocks.py:110-117)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        assert groups == 1
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = O.as_triple(filter_size, dimension)
        self.filter_stride = O.as_triple(filter_stride, dimension)
        self.filter_volume = int(np.prod(self.filter_size))
        self.weight = _init_weight(self.filter_volume, nIn, nOut)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn
        md, sp = input.metadata, _sp(input)
        out_sp, rules = md.get_strided(sp, self.filter_size, self.filter_stride)
        out = SparseConvNetTensor(metadata=md, spatial_size=torch.LongTensor(list(out_sp)))
        out.features = _ConvFn.apply(input.features, self.weight, self.bias, rules, md.levels[out_sp].shape[0])
        return out


class Deconvolution(nn.Module):
    """(dimension, nIn, nOut, filter_size, filter_stride, bias)  (sparse_building_blocks.py:207-213)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = O.as_triple(filter_size, dimension)
        self.filter_stride = O.as_triple(filter_stride, dimension)
        self.filter_volume = int(np.prod(self.filter_size))
        self.weight = _init_weight(self.filter_volume, nIn, nOut)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, input):
        md, sp = input.metadata, _sp(input)
        fine = tuple((sp[a] - 1) * self.filter_stride[a] + self.filter_size[a] for a in range(3))
        assert fine in md.levels, "Deconvolution needs the fine grid to exist in the metadata (App. A.7)"
        _, rules = md.get_strided(fine, self.filter_size, self.filter_stride)
        out = SparseConvNetTensor(metadata=md, spatial_size=torch.LongTensor(list(fine)))
        out.features = _ConvFn.apply(input.features, self.weight, self.bias, O.swap_rules(rules),
                                     md.levels[fine].shape[0])
        return out


class AveragePooling(nn.Module):
    """(dimension, pool_size, pool_stride, nFeaturesToDrop=0)  (sparse_building_blocks.py:150-154)."""

    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        super().__init__()
        self.dimension = dimension
        self.pool_size = O.as_triple(pool_size, dimension)
        self.pool_stride = O.as_triple(pool_stride, dimension)
        self.pool_volume = int(np.prod(self.pool_size))
        self.nFeaturesToDrop = int(nFeaturesToDrop)

    def forward(self, input):
        md, sp = input.metadata, _sp(input)
        out_sp, rules = md.get_strided(sp, self.pool_size, self.pool_stride)
        out = SparseConvNetTensor(metadata=md, spatial_size=torch.LongTensor(list(out_sp)))
        out.features = _AvgPoolFn.apply(input.features, rules, md.levels[out_sp].shape[0], self.pool_volume,
                                        self.nFeaturesToDrop)
        return out


class BatchNormalization(nn.Module):
    """(nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1)  (sparse_building_blocks.py:39,122)."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness = nPlanes, eps, momentum, affine, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = nn.Parameter(torch.ones(nPlanes))
            self.bias = nn.Parameter(torch.zeros(nPlanes))
        else:
            self.weight = self.bias = None

    def forward(self, input):
        x = input.features
        assert x.nelement() == 0 or x.size(1) == self.nPlanes
        training = self.training

        def run(leak=float(self.leakiness)):
            return _BNFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, training, self.eps,
                               self.momentum, leak)
        if NUMERICS["fuse"] and float(self.leakiness) == 1.0:
            return SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size, pending=(run, run))
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = run()
        return out


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)


class LeakyReLU(nn.Module):
    def __init__(self, leak=1.0 / 3.0):
        super().__init__()
        self.leak = leak

    def forward(self, input):
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        pend, input._pending = input._pending, None
        if pend is not None:
            out.features = pend[1](float(self.leak))
        else:
            out.features = _LeakyFn.apply(input.features, self.leak)
        return out


class ReLU(LeakyReLU):
    def __init__(self):
        super().__init__(0.0)


class Tanh(nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = torch.tanh(input.features)
        return out


class Sigmoid(nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(metadata=input.metadata, spatial_size=input.spatial_size)
        out.features = torch.sigmoid(input.features)
        return out


class Identity(nn.Module):
    def forward(self, input):
        return input


class AddTable(nn.Module):
    def forward(self, input):
        if NUMERICS["fuse"] and len(input) == 2:
            a, b = input[0].features, input[1].features
            return SparseConvNetTensor(metadata=input[0].metadata, spatial_size=input[0].spatial_size,
                                       pending=(lambda: _AddFn.apply(a, b), lambda leak: _AddLeakyFn.apply(a, b, leak)))
        out = SparseConvNetTensor(metadata=input[0].metadata, spatial_size=input[0].spatial_size)
        feats = input[0].features
        for t in input[1:]:
            feats = _AddFn.apply(feats, t.features)
        out.features = feats
        return out


class SparseToDense(nn.Module):
    """(dimension, nPlanes)  (src/networks/resnet.py:123-125)."""

    def __init__(self, dimension, nPlanes):
        super().__init__()
        self.dimension, self.nPlanes = dimension, nPlanes

    def forward(self, input):
        md, sp = input.metadata, _sp(input)
        return _S2DFn.apply(input.features, md.levels[sp], sp, md.batch_size)


class Sequential(nn.Sequential):
    pass
