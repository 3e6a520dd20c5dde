import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _tensor_core_mode_for_gpu_tests(request):
    """The package defaults to SparseConvNet's fp32 numerics; the GPU tests were written against the tensor-core mode
    ("bf16", what bench.py times) and switch explicitly where they want another one.  Start each from "bf16"."""
    if "gpu" in request.keywords:
        try:
            import torch
            if torch.cuda.is_available():
                import sparseconvnet as scn
                scn.set_precision("bf16")
        except Exception:
            pass
    yield
