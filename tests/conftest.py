import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CKPT = os.path.join(ROOT, "ckpt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in "fiu"
                    else str(z[k])) for k in z.files}


def load_params(ckpt_name):
    """checkpoint['params'] of ckpt/WaveMamba_<name>.pth (keys keep their prefix)."""
    return torch.load(os.path.join(CKPT, f"WaveMamba_{ckpt_name}.pth"), map_location="cpu")["params"]


@pytest.fixture(scope="session")
def params_cache():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_params(name)
        return cache[name]
    return get
