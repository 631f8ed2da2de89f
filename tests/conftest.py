"""pytest configuration: registers the ``gpu`` marker and puts the repository root on ``sys.path``.

``-m "not gpu"`` runs on any machine (oracle vs golden vectors, host logic, C-ABI symbol check, gloo tests);
``-m gpu`` needs a B200 and calls the CUDA path through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA sm_100 device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def built_library():
    """Build (or reuse) the in-tree shared library; nvcc cross-compiles without a GPU."""
    from pylrbms_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def handle(built_library):
    import torch
    if not torch.cuda.is_available():
        pytest.fail('gpu-marked test selected on a machine without CUDA: the product path has no CPU fallback')
    from pylrbms_b200._lib import Handle
    return Handle.get(0)
