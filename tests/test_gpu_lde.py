"""GPU parity: LDE through the C-ABI against the C oracle (itself pinned to oracle/ntt.py) -- bit-exact."""
import numpy as np
import pytest

from conftest import rand_felts
from oracle import clib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n,n_cols,log_blowup", [(0, 1, 1), (3, 2, 3), (9, 3, 2), (12, 5, 3), (16, 3, 3), (20, 1, 1)])
def test_lde_vs_c_oracle(ctx, log_n, n_cols, log_blowup):
    tr = rand_felts(n_cols << log_n, 700 + log_n)
    got = ctx.lde(tr, log_n, n_cols, log_blowup)
    want = clib.lde(tr, log_n, n_cols, log_blowup)
    assert np.array_equal(got, want)


def test_lde_other_offset(ctx):
    tr = rand_felts(2 << 10, 77)
    off = 0x1234567890abcdef1234567890abcdef
    assert np.array_equal(ctx.lde(tr, 10, 2, 2, offset=off), clib.lde(tr, 10, 2, 2, offset=off))


def test_lde_first_coset_interpolates(ctx):
    """Size-independent property: with offset 1 the coset j=0 reproduces the trace itself."""
    log_n, n_cols = 18, 2
    tr = rand_felts(n_cols << log_n, 78)
    out = ctx.lde(tr, log_n, n_cols, 1, offset=1)
    assert np.array_equal(out[: n_cols << log_n], tr)
