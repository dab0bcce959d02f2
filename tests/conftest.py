import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from xm_helpers import GOLDEN, load_golden_tables  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: skip them cleanly where there is none (a plain `pytest tests` on a
    CPU box then reports skips, not errors)."""
    try:
        import torch

        have_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_cuda = False
    if have_cuda:
        return  # on a GPU box nothing is skipped: a missing library must fail loudly, not pass silently
    skip = pytest.mark.skip(reason="gpu test: no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def tables_default():
    return load_golden_tables("default")[0]


@pytest.fixture(scope="session")
def tables_small():
    return load_golden_tables("small")[0]
