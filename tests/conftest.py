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
