import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "few-shot-music-generation_b200"
for p in (str(ROOT), str(PKG), str(PKG / "src")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """libfsmg.so, built in-tree (nvcc cross-compiles on the CPU box)."""
    import __graft_entry__ as ge
    ge.build()
    return ge.LIB


def load_golden(name):
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_golden
    cfg = dict(make_golden.CASES[name][0], name="lstm_baseline", lr=5e-3, n_decay=10000, max_grad_norm=5)
    blob = dict(np.load(ROOT / "tests" / "golden" / f"{name}.npz"))
    return cfg, blob
