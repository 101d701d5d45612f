"""pytest configuration: `gpu` marker, golden-vector loader, repo root on sys.path."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu); everything else runs on CPU")


def load_golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix: str) -> list[str]:
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.fixture(scope="session")
def cuda_device():
    """GPU tests must run the CUDA path for real: no skipping, no fallback."""
    import torch
    assert torch.cuda.is_available(), "a test marked `gpu` was run without a CUDA device"
    from snag_b200 import _lib
    _lib.call("snag_device_check")
    return torch.device("cuda:0")
