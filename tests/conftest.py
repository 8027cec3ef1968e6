import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "crypto_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ctx():
    """The CUDA context; GPU tests must run on the native library, never on a fallback."""
    import stark_perpetual_b200 as spg
    return spg.get_context(0)


def rand_felts(n, seed):
    """(n, 4) uint64 canonical felts, numpy PCG64, rejection-sampled below p."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    p3 = 0x0800000000000011
    out = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    bad = out[:, 3] >= np.uint64(p3)      # conservative: reject anything with top limb >= p3
    while bad.any():
        k = int(bad.sum())
        out[bad] = rng.integers(0, 2**64, size=(k, 4), dtype=np.uint64)
        out[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
        bad = out[:, 3] >= np.uint64(p3)
    return out
