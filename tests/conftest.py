"""pytest config: registers the `gpu` marker; puts tests/ on sys.path so helper modules import by name."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = Path(__file__).resolve().parent / "golden"
    return {n: np.load(d / f"{n}.npz") for n in ("mul_mat", "ops", "flash_attn")}
