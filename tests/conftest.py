import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build every native artefact once per session (idempotent `make`)."""
    import __graft_entry__ as g
    g.build()
    return True


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a host without one they are skipped instead of failing, so a plain
    `pytest tests` is green on CPU-only CI.  The probe goes through the library's own entry point; a MISSING library
    is not a reason to skip (the product has no fallback) -- the import error surfaces in the tests themselves."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        from fractalshark_b200.gpu_renderer import GPURenderer
        working = bool(GPURenderer.TestCudaIsWorking())
    except Exception:
        return
    if not working:
        skip = pytest.mark.skip(reason="no working CUDA device on this host (fs_test_cuda_is_working() == 0)")
        for it in gpu_items:
            it.add_marker(skip)
