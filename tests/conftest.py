import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU checker: the plain-C restatement is always built; the reference build is used when present."""
    from oracle import pyoracle
    pyoracle.build(port=True, ref=os.path.isdir("/root/reference"))
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fsgm_b200 import api
    c = api.Context(0)
    c.use_torch_stream()
    yield c
    c.close()
