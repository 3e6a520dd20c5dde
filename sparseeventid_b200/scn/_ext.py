"""Loader of the thin PyTorch C++ layer (build_torch/scn_b200_torch.so, built in-tree by
``python -m sparseeventid_b200.build --torch`` / ``__graft_entry__.build()``).

It holds the autograd functions of the hot modules in C++ (csrc_torch/scn_torch.cpp) so a module forward/backward
costs one native call.  If it has not been built the Python functions in functional.py run instead -- the same
kernels through the same C ABI, only with more interpreter time per module.  ``SCN_B200_TORCH_EXT=0`` disables it.
"""
from __future__ import annotations

import ctypes
import importlib.machinery
import importlib.util
import os

from .. import _lib
from ..build import TORCH_EXT_NAME, TORCH_EXT_SO

_mod = None
_tried = False


def get():
    """The extension module, or None."""
    global _mod, _tried
    if _tried:
        return _mod
    _tried = True
    if os.environ.get("SCN_B200_TORCH_EXT", "1") in ("0", "false", "False") or not os.path.exists(TORCH_EXT_SO):
        return None
    import torch  # noqa: F401  (libtorch must be mapped before the extension)
    _lib.load()
    try:
        ctypes.CDLL(_lib.LIB_PATH, mode=ctypes.RTLD_GLOBAL)   # resolves the extension's NEEDED libscn_b200.so
        loader = importlib.machinery.ExtensionFileLoader(TORCH_EXT_NAME, TORCH_EXT_SO)
        spec = importlib.util.spec_from_loader(TORCH_EXT_NAME, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _mod = mod
    except (ImportError, OSError) as e:       # e.g. built against another torch: the Python functions take over
        import warnings
        warnings.warn(f"scn_b200_torch.so could not be loaded ({e}); using the Python autograd functions "
                      "(same CUDA kernels through the same C ABI)")
        _mod = None
    return _mod
