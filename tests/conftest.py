import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    if os.environ.get("KFP16_FORCE_NO_GPU"):
        return False
    try:
        import ctypes

        cudart = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch

            return torch.cuda.is_available()
        except Exception:
            return False


HAVE_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAVE_GPU
    if HAVE_GPU is None:
        HAVE_GPU = _have_gpu()
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    from kaldi_fp16_b200 import _lib

    return _lib.load()


@pytest.fixture(scope="session")
def handle(lib):
    from kaldi_fp16_b200 import gpu

    gpu.Init(0)
    h = gpu.NewHandle()
    yield h
    h.Destroy()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own operator library compiled unmodified for sm_100a (oracle/Makefile)."""
    from tests.refbind import load_ref

    lib = load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libkaldi_fp16_ref.so not built")
    return lib
