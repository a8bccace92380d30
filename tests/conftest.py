"""pytest configuration: markers, paths, shared fixtures."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

TESTS = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(TESTS)
GOLDEN = os.path.join(TESTS, "golden")
for p in (REPO, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference")
    config.addinivalue_line("markers", "slow: large-N property tests")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load


REFERENCE_ROOT = os.environ.get("SMC_REFERENCE_ROOT", "/root/reference")
needs_reference = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")),
    reason="reference tree not present (only in the build container)")
