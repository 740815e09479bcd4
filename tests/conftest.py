import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs >1 CUDA device")
    import torch
    if not torch.cuda.is_available():
        # CPU suite: a model tensor that the converted checkpoint does not provide is an error (on the GPU box it stays a warning)
        os.environ.setdefault("B200_STRICT_LOAD", "1")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    import torch
    torch.manual_seed(0)
    yield
