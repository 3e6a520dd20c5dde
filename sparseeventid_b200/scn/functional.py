"""autograd.Function wrappers: the host-side mirror of SCN's ``X_updateOutput`` / ``X_backward``
pairs.  Forward and backward both run on the C ABI; nothing is computed by torch.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import config, ops


def _w3(weight):
    """SCN stores [K, groups=1, Cin, Cout] (newer) or [K, Cin, Cout] (2018-19); kernels take 3-D."""
    return weight.detach().reshape(weight.shape[0], weight.shape[-2], weight.shape[-1]).contiguous().float()


class ConvFn(Function):
    """out[o] = bias + sum_k x[nbr_fwd[k][o]] @ W[k].

    mirror=True  (submanifold): dgrad uses the same table with B_k = W[K-1-k]^T.
    mirror=False (strided / deconvolution): dgrad uses nbr_bwd (the transposed table) with W[k]^T.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, nbr_fwd, nbr_bwd, n_out_rows, mirror):
        x = x.contiguous()
        prec = config.precision_code()
        w3 = _w3(weight)
        K, cin, cout = w3.shape
        bprep = ops.prep_weights(w3, False, False, prec, config.feature_dtype())
        b = bias.detach().float().contiguous() if bias is not None else None
        out = ops.conv_forward(x, nbr_fwd, n_out_rows, cin, cout, bprep, b, prec, config.feature_dtype())
        ctx.save_for_backward(x, weight)
        ctx.nbr_fwd, ctx.nbr_bwd, ctx.mirror, ctx.prec = nbr_fwd, nbr_bwd, mirror, prec
        ctx.has_bias = bias is not None
        ctx.n_out_rows = n_out_rows
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight = ctx.saved_tensors
        dout = dout.contiguous()
        w3 = _w3(weight)
        K, cin, cout = w3.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            bt = ops.prep_weights(w3, True, ctx.mirror, ctx.prec, x.dtype)
            dx = ops.conv_forward(dout, ctx.nbr_bwd, x.shape[0], cout, cin, bt, None, ctx.prec, x.dtype,
                                  kind="conv_dgrad")
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(x, dout, ctx.nbr_fwd, ctx.n_out_rows, cin, cout, ctx.prec).view_as(weight)
            dw = dw.to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.col_sum(dout)
        return dx, dw, db, None, None, None, None


class BatchNormFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, eps, momentum, leak):
        x = x.contiguous()
        g = weight.detach().float().contiguous() if weight is not None else None
        b = bias.detach().float().contiguous() if bias is not None else None
        out, mean, invstd = ops.bn_forward(x, g, b, running_mean, running_var, training, eps, momentum, leak)
        ctx.save_for_backward(x, g, b, mean, invstd)
        ctx.training, ctx.leak, ctx.affine = training, leak, weight is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, g, b, mean, invstd = ctx.saved_tensors
        dx, dg, db = ops.bn_backward(x, dout.contiguous(), g, b, mean, invstd, ctx.training, ctx.leak)
        if not ctx.affine:
            dg = db = None
        return dx, dg, db, None, None, None, None, None, None


class LeakyReLUFn(Function):
    @staticmethod
    def forward(ctx, x, leak):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.leak = leak
        return ops.leaky_forward(x, leak)

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        return ops.leaky_backward(x, dout.contiguous(), ctx.leak), None


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add_forward(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, dout):
        return dout, dout


class AddLeakyFn(Function):
    """out = leaky(a + b): AddTable followed by LeakyReLU in one kernel (reference ResidualBlock tail,
    src/networks/sparse_building_blocks.py:96-98)."""

    @staticmethod
    def forward(ctx, a, b, leak):
        out = ops.add_forward(a.contiguous(), b.contiguous(), leak)
        ctx.save_for_backward(out)
        ctx.leak = leak
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        # sign(out) == sign(a + b) for leak > 0; for leak == 0 (ReLU) out > 0 <=> a + b > 0 as well
        d = ops.leaky_backward(out, dout.contiguous(), ctx.leak)
        return d, d, None


class InputLayerFn(Function):
    @staticmethod
    def forward(ctx, feats, rows, n_active, mode):
        ctx.rows, ctx.mode, ctx.in_dtype = rows, mode, feats.dtype
        return ops.input_layer_forward(feats.detach().float().contiguous(), rows, n_active, mode)

    @staticmethod
    def backward(ctx, dout):
        if ctx.mode not in (0, 3):
            raise NotImplementedError("InputLayer backward is implemented for modes 0 and 3")
        return ops.rows_gather(dout.contiguous(), ctx.rows, ctx.in_dtype if ctx.in_dtype in (torch.float32, torch.bfloat16) else torch.float32), None, None, None


class OutputLayerFn(Function):
    @staticmethod
    def forward(ctx, feats, rows):
        ctx.rows, ctx.n, ctx.dtype = rows, feats.shape[0], feats.dtype
        return ops.rows_gather(feats.contiguous(), rows, feats.dtype)

    @staticmethod
    def backward(ctx, dout):
        acc = ops.rows_scatter_add(dout.contiguous(), ctx.rows, ctx.n)
        return ops.convert(acc, ctx.dtype), None


class SparseToDenseFn(Function):
    @staticmethod
    def forward(ctx, x, keys, batch, spatial):
        x = x.contiguous()
        ctx.keys, ctx.batch, ctx.spatial = keys, batch, spatial
        ctx.n, ctx.c, ctx.dtype = x.shape[0], x.shape[1], x.dtype
        return ops.sparse_to_dense_forward(x, keys, batch, spatial)

    @staticmethod
    def backward(ctx, ddense):
        dx = ops.sparse_to_dense_backward(ddense.contiguous().float(), ctx.keys, ctx.n, ctx.c, ctx.batch, ctx.spatial,
                                          ctx.dtype)
        return dx, None, None, None
