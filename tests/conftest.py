import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the reference (run unmodified by tests/test_oracle_vs_reference.py) calls numpy.trapz, which
    # numpy >= 2.0 deprecates, and divides by zero by design at masked points (kernels.py:67)
    config.addinivalue_line("filterwarnings", "ignore:`trapz` is deprecated:DeprecationWarning")
    config.addinivalue_line("filterwarnings", "ignore::RuntimeWarning:agnpy")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
