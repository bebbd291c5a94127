import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Read-only view of one tests/golden/*.npz fixture; `g.t("case.key")` -> torch tensor."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)

    def has(self, key):
        return key in self.z.files

    def a(self, key):
        return self.z[key]

    def t(self, key, device="cpu"):
        arr = self.z[key]
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if t.dtype == torch.int32:
            t = t.long()
        return t.to(device)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return get


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def ref_module():
    """The reference's own CPU extension (oracle/_ref), or None when it is not present."""
    from oracle import build_ref

    if not build_ref.available():
        try:
            build_ref.build(verbose=False)
        except Exception:
            return None
    if not build_ref.available():
        return None
    return build_ref.load()
