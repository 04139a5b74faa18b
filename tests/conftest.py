import os
import sys

import numpy as np
import pytest

# the five .onnx files cannot be downloaded offline: the tests opt in to seeded random weights of the same architectures
os.environ.setdefault("B2F_SYNTHETIC_WEIGHTS", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes tens of seconds on the CPU box")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def ref():
    """The reference's own modules, imported verbatim; None on the GPU box (no /root/reference there)."""
    from oracle import ref_loader
    return ref_loader.load()
