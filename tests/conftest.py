import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box with -m gpu")


@pytest.fixture(scope="session")
def salun_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from unlearn_saliency_b200.tail import SalunContext
    ctx = SalunContext(0)
    yield ctx
    ctx.close()
