import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "quantized-gemm-for-transformer-inference_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def qg():
    """The product package (hyphenated directory name, so imported through importlib)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def oracle():
    import oracle as o

    o.build()
    return o


def make_edge_matrix(rng: np.random.Generator, rows: int, cols: int, scale: float = 1.0) -> np.ndarray:
    """Random matrix whose first rows exercise the reference's quantizer quirks."""
    X = (rng.random((rows, cols), dtype=np.float32) * 2 - 1) * scale
    r = 0

    def take():
        nonlocal r
        i = r
        r += 1
        return i if i < rows else None

    i = take()
    if i is not None:  # all-zero vector: absmax 0 -> scale inf -> 0*inf = NaN -> code 0
        X[i, :] = 0.0
    i = take()
    if i is not None and cols > 1:  # largest magnitude is a NEGATIVE first element: scale too small, codes wrap
        X[i, 0] = -3.0 * scale
    i = take()
    if i is not None and cols > 1:  # negative first element, every later entry zero: +-0 tie-break path
        X[i, :] = 0.0
        X[i, 0] = -0.5
    i = take()
    if i is not None and cols > 2:  # same with a negative zero as the first later entry
        X[i, :] = 0.0
        X[i, 0] = -0.5
        X[i, 1] = -0.0
    i = take()
    if i is not None and cols > 3:  # NaN and inf inside the vector
        X[i, 2] = np.nan
        X[i, 3] = np.inf
    i = take()
    if i is not None:  # NaN as the first element
        X[i, 0] = np.nan
    i = take()
    if i is not None:  # -0.0 first element with zeros behind it
        X[i, :] = 0.0
        X[i, 0] = -0.0
    i = take()
    if i is not None and cols > 1:  # positive first element is the max
        X[i, 0] = 5.0 * scale
    return X
