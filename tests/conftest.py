import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # placements found by the tests go to a scratch cache directory, not to the user's ~/.cache
    import tempfile
    os.environ.setdefault("LDPC_B200_CACHE_DIR", tempfile.mkdtemp(prefix="ldpc_b200_cache_"))


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library and the oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import oracle as O
    return O
