"""ctypes loader of the plain-C oracle (oracle/scn_oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "build", "liboracle_c.so")


def build() -> str:
    src = os.path.join(HERE, "scn_oracle_c.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "build/liboracle_c.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        p, i64, i = C.c_void_p, C.c_int64, C.c_int
        l.oc_input_rules.restype = i64
        l.oc_input_rules.argtypes = [p, i64, p, p]
        l.oc_subm_rules.restype = i
        l.oc_subm_rules.argtypes = [p, i64, i, i, i, p, p, p]
        l.oc_strided_rules.restype = i64
        l.oc_strided_rules.argtypes = [p, i64, i, i, i, p, p, p]
        l.oc_conv_forward.restype = None
        l.oc_conv_forward.argtypes = [p, p, p, p, p, p, i, i, i, i64, p]
        _lib = l
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def input_rules(coords):
    coords = np.ascontiguousarray(coords, dtype=np.int64)
    n = coords.shape[0]
    rows = np.empty((n,), np.int64)
    active = np.empty((max(n, 1), 4), np.int64)
    m = lib().oc_input_rules(_ptr(coords), n, _ptr(rows), _ptr(active))
    return rows, active[:m]


def subm_rules(coords, filt):
    coords = np.ascontiguousarray(coords, dtype=np.int64)
    n = coords.shape[0]
    K = filt[0] * filt[1] * filt[2]
    pin = np.empty((max(K * n, 1),), np.int32)
    pout = np.empty((max(K * n, 1),), np.int32)
    offs = np.empty((K + 1,), np.int64)
    rc = lib().oc_subm_rules(_ptr(coords), n, filt[0], filt[1], filt[2], _ptr(pin), _ptr(pout), _ptr(offs))
    assert rc == 0
    return [np.stack([pin[offs[k]:offs[k + 1]], pout[offs[k]:offs[k + 1]]], 1).astype(np.int64) for k in range(K)], \
        (pin, pout, offs)


def strided_rules(coords, stride):
    coords = np.ascontiguousarray(coords, dtype=np.int64)
    n = coords.shape[0]
    out_coords = np.empty((max(n, 1), 4), np.int64)
    out_row = np.empty((max(n, 1),), np.int32)
    off = np.empty((max(n, 1),), np.int32)
    m = lib().oc_strided_rules(_ptr(coords), n, stride[0], stride[1], stride[2], _ptr(out_coords), _ptr(out_row), _ptr(off))
    K = stride[0] * stride[1] * stride[2]
    rules = []
    for k in range(K):
        sel = np.nonzero(off[:n] == k)[0]
        pr = np.stack([sel, out_row[:n][sel]], 1).astype(np.int64)
        rules.append(pr[np.lexsort((pr[:, 0], pr[:, 1]))] if pr.shape[0] else pr.reshape(0, 2))
    return out_coords[:m], rules


def conv_forward(x, W, bias, raw, n_out):
    pin, pout, offs = raw
    x = np.ascontiguousarray(x, dtype=np.float64)
    W = np.ascontiguousarray(W, dtype=np.float64)
    K, cin, cout = W.shape
    out = np.empty((n_out, cout), np.float64)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float64)
    lib().oc_conv_forward(_ptr(x), _ptr(W), None if b is None else _ptr(b), _ptr(pin), _ptr(pout), _ptr(offs), K, cin, cout,
                          n_out, _ptr(out))
    return out
