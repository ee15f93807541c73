import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def ref_state_dict(golden):
    import torch
    g = golden("model_seed0.npz")
    return {k: torch.from_numpy(g[k].copy()) for k in g.files}


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def gpu_model(cuda_dev, ref_state_dict):
    import ertdiff_b200 as eb
    m = eb.ConditionalDiffusionModel(29, 128)
    m.load_state_dict(ref_state_dict)
    return m.to(cuda_dev).eval()


def parity_error(got, want, rtol, atol):
    """max over elements of |got - want| / (atol + rtol*|want|): <= 1 means every element is within
    ``atol + rtol*|want|`` (the numpy.allclose criterion, element by element)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if got.size == 0:
        return 0.0
    return float((np.abs(got - want) / (atol + rtol * np.abs(want))).max())


def assert_close(got, want, rtol, atol, what=""):
    """Element-wise ``|got - want| <= atol + rtol*|want|``; the message carries the measured errors."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    ratio = parity_error(got, want, rtol, atol)
    if not ratio <= 1.0:
        d = np.abs(got - want)
        raise AssertionError(f"{what}: worst element at {ratio:.3g}x its tolerance (rtol={rtol:g}, atol={atol:g}); "
                             f"max|d|={d.max():.3e}, max|want|={np.abs(want).max():.3e}, "
                             f"max|d|/|want|={float((d / np.maximum(np.abs(want), 1e-30)).max()):.3e}")
