import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden" / "tiny_ref.npz"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN, allow_pickle=False))


@pytest.fixture(scope="session")
def variants():
    """Golden vectors of the f-4 variants (oracle/make_golden_variants.py, live reference classes)."""
    return dict(np.load(REPO / "tests" / "golden" / "variants_ref.npz", allow_pickle=False))


@pytest.fixture(scope="session")
def tiny_lists(golden):
    """(train lists in file order, test dict) rebuilt from the golden edge arrays."""
    n = int(golden["n_users"])
    train = [[] for _ in range(n)]
    for u, i in zip(golden["train_user"].tolist(), golden["train_item"].tolist()):
        train[u].append(i)
    test = {}
    for u, i in zip(golden["test_user"].tolist(), golden["test_item"].tolist()):
        test.setdefault(u, []).append(i)
    return [np.array(t, dtype=np.int64) for t in train], test
