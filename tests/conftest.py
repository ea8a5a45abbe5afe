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
    """The reference's own modules, importable only in the build container (read-only mount)."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("/root/reference not present (GPU box)")
    import types
    for n in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.transform"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].rc = lambda *a, **k: None
    sys.modules["matplotlib.pyplot"].figure = lambda *a, **k: None
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    sys.modules["skimage.transform"].resize = None
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import models.mygannet as mg
    import models.spatiotempconv as stc
    import models.convlstm as cl
    import lib.utils as lu
    return types.SimpleNamespace(mygannet=mg, spatiotempconv=stc, convlstm=cl, utils=lu)
