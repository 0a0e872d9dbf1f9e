import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box with -m gpu")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only box: GPU-marked tests are skipped, not failed."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def salun_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from unlearn_saliency_b200.tail import SalunContext
    ctx = SalunContext(0)
    yield ctx
    ctx.close()
