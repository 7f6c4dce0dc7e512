import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ppx():
    return importlib.import_module("pairwise-perturbation_b200")


@pytest.fixture(scope="session")
def ctx(ppx):
    c = ppx.Ctx(0, workspace_bytes=512 << 20)
    yield c
    c.close()
