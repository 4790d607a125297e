import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    torch.set_num_threads(max(1, min(4, os.cpu_count() or 1)))


def load_weights(bands):
    z = np.load(os.path.join(GOLDEN, f"weights_b{bands}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def weights4():
    return load_weights(4)


@pytest.fixture(scope="session")
def weights8():
    return load_weights(8)
