import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped, not failed, on a host that has no CUDA device (or no built library): the product
    has no CPU path to fall back to."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    reason = None
    try:
        from metagraph_b200 import _lib
        if _lib.load_library().mgb_device_count() < 1:
            reason = "no CUDA device"
    except (ImportError, OSError) as e:
        reason = "libmgb.so not available: %s" % e
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    import oracle_lib
    oracle_lib.build_oracle()
