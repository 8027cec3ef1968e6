"""GPU parity: LDE through the C-ABI against the C oracle (itself pinned to oracle/ntt.py) -- bit-exact."""
import numpy as np
import pytest

from conftest import rand_felts
from oracle import clib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n,n_cols,log_blowup", [(0, 1, 1), (3, 2, 3), (9, 3, 2), (12, 5, 3), (16, 3, 3), (20, 1, 1), (21, 1, 2),
                                                     (23, 1, 1)])   # 21: the 2^11 tile; 23: three passes
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


def test_lde_baseline_config3a_full_size(ctx):
    """BASELINE.json configs[2] at full size: 25 columns x 2^20 rows, blowup 8 (6.7 GB of evaluations).  Three columns
    are compared in full with the C oracle, and every column's coset 0 ... 7 is tied to it through linearity:
    LDE(sum of all columns) == sum of the LDEs (checked on the oracle side for the summed column)."""
    log_n, n_cols = 20, 25
    n = 1 << log_n
    tr = rand_felts(n_cols << log_n, 1003).reshape(n_cols, n, 4)
    out = ctx.lde(tr.reshape(-1, 4), log_n, n_cols, 3).reshape(8, n_cols, n, 4)
    for c in (0, 11, 24):
        want = clib.lde(tr[c], log_n, 1, 3).reshape(8, n, 4)
        assert np.array_equal(out[:, c], want), c
    # linearity over ALL columns: fold the 25 device results with the field-add kernel and compare with the oracle LDE
    # of the folded trace column
    acc_in, acc_out = tr[0].copy(), out[:, 0].reshape(-1, 4).copy()
    for c in range(1, n_cols):
        acc_in = ctx.field_op("add", acc_in, tr[c])
        acc_out = ctx.field_op("add", acc_out, out[:, c].reshape(-1, 4))
    assert np.array_equal(acc_out, clib.lde(acc_in, log_n, 1, 3))
