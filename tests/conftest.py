import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "reference-qvm_b200")
for p in (REPO, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

import referenceqvm  # noqa: E402,F401  (also puts the pyquil stand-in on sys.path when pyquil is absent)

assert not referenceqvm.__file__.startswith("/root/reference"), "tests must run against this repo's package"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="module")
def qvm():
    from referenceqvm.api import QVMConnection
    return QVMConnection(type_trans="wavefunction")


@pytest.fixture(scope="module")
def qvm_unitary():
    from referenceqvm.api import QVMConnection
    return QVMConnection(type_trans="unitary")
