import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0:
            return n.value
        return 0
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """gpu tests need a device: skip them (instead of failing in the product's no-device error) when a
    run without `-m "not gpu"` lands on a machine that has none."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle():
    from oracle import md_oracle
    md_oracle.lib()
    return md_oracle
