import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when selected without a device: a silent skip would hide a
    # missing CUDA path.  They are simply deselected by `-m "not gpu"` on the CPU tier.
    pass
