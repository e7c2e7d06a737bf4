import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def slot_model_path():
    from av_aloha_b200 import model_io
    return model_io.model_path("slot_insertion", 3)


@pytest.fixture(scope="session", autouse=True)
def _native_pieces_built():
    """libavsim.so (product) and the oracle / emulation harness (checkers) are build artefacts, not tracked files: make
    sure they exist before any test loads them (nvcc cross-compiles without a GPU; a few tens of seconds when stale)."""
    import __graft_entry__ as g
    try:
        g.build_cuda()
    except Exception as e:  # noqa: BLE001 - e.g. no nvcc on the box: the prebuilt .so that travelled with the snapshot is used
        if not os.path.exists(g.LIB):
            raise RuntimeError(f"libavsim.so is missing and could not be built: {e}")
    yield
