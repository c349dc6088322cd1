import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    """GPU tests are selected with `-m gpu`; if they are collected on a machine without a GPU
    (plain `pytest tests/`), skip rather than fail."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _built_libraries():
    """Both libraries are built in-tree before the session (no-ops when up to date)."""
    from oracle import oracle as orc
    orc.build()
    from wurm_b200 import build as wb
    wb.build()
