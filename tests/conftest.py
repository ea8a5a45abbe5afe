import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def reference_modules():
    """The reference's own modules: from the read-only mount in the build container, else from the copy staged
    under oracle/_ref (oracle/make_ref.py; shipped to the GPU box)."""
    from oracle import make_ref
    if make_ref.ref_root() is None:
        pytest.skip("the reference is neither at /root/reference nor staged under oracle/_ref")
    return make_ref.import_ref()
