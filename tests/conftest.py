import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle.load_oracle()


@pytest.fixture(scope="session")
def ref():
    import _oracle
    r = _oracle.load_ref()
    if r is None:
        pytest.skip("oracle/_ref/libnvpyr_ref.so not built (reference tree absent)")
    return r


@pytest.fixture(scope="session")
def emu():
    import _oracle
    e = _oracle.load_emu()
    if e is None:
        pytest.skip("oracle/_ref/libnvpyr_glsl_emu.so not built (reference tree absent)")
    return e


@pytest.fixture(scope="session")
def nv():
    import vk_compute_mipmaps_b200 as m
    return m


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch
