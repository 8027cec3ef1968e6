"""GPU: empty, minimal and extreme inputs through every batched entry point (the domain's edge cases: empty batches,
single elements, field-boundary values, collisions), plus size-independent properties at full size."""
import numpy as np
import pytest

from conftest import rand_felts
from oracle import clib
from oracle.params import FIELD_PRIME as P
from oracle.pedersen import pedersen_hash as opedersen
from stark_perpetual_b200._lib import NTT_NAT_TO_REV, NTT_REV_TO_NAT, ints_to_limbs, limbs_to_ints

pytestmark = pytest.mark.gpu
E4 = np.empty((0, 4), dtype=np.uint64)


def test_empty_batches(ctx):
    out, st = ctx.pedersen_hash2(E4, E4)
    assert out.shape == (0, 4) and st.shape == (0,)
    out, st = ctx.pedersen_chain(E4, 3)
    assert out.shape == (0, 4)
    assert ctx.ecdsa_verify(E4, E4, E4, E4).shape == (0,)
    out, st = ctx.private_to_stark_key(E4)
    assert out.shape == (0, 4)
    assert ctx.field_op("mul", E4, E4).shape == (0, 4)
    assert ctx.ntt(E4, 5).shape == (0, 4)
    empty_orders = {"asset_id_synthetic": E4, "asset_id_collateral": E4, "asset_id_fee": E4,
                    "is_buying_synthetic": np.empty(0, np.uint8), "amount_synthetic": np.empty(0, np.uint64),
                    "amount_collateral": np.empty(0, np.uint64), "max_amount_fee": np.empty(0, np.uint64),
                    "position_id": np.empty(0, np.uint64), "nonce": np.empty(0, np.uint32),
                    "expiration_timestamp": np.empty(0, np.uint32)}
    out, st = ctx.limit_order_msg(empty_orders)
    assert out.shape == (0, 4) and st.shape == (0,)
    r, s, st = ctx.sign(E4, E4)
    assert r.shape == (0, 4) and s.shape == (0, 4) and st.shape == (0,)
    out, st = ctx.message_hash("price", [E4, E4], [np.empty(0, np.uint64)] * 2)
    assert out.shape == (0, 4) and st.shape == (0,)


def test_single_element_and_size_one_transforms(ctx):
    x = ints_to_limbs([P - 1])
    assert limbs_to_ints(ctx.ntt(x, 0)) == [P - 1]                       # size-1 transform is the identity
    assert limbs_to_ints(ctx.ntt(x, 0, inverse=True)) == [P - 1]
    pair = ints_to_limbs([5, P - 1])
    assert limbs_to_ints(ctx.ntt(pair, 1, order=2)) == [(5 + P - 1) % P, (5 - (P - 1)) % P]
    out, st = ctx.pedersen_hash2(ints_to_limbs([0]), ints_to_limbs([0]))
    assert st[0] == 0 and limbs_to_ints(out)[0] == opedersen(0, 0)
    lde = ctx.lde(ints_to_limbs([7] * 8), 3, 1, 3)                        # a constant column extends to the constant
    assert limbs_to_ints(lde) == [7] * 64


def test_pedersen_boundary_values(ctx):
    vals = [0, 1, P - 1, P - 2, 2**248 - 1, 2**248, 2**251, 2**251 + 17 * 2**192]
    xs = [a for a in vals for _b in vals]
    ys = [b for _a in vals for b in vals]
    out, st = ctx.pedersen_hash2(ints_to_limbs(xs), ints_to_limbs(ys))
    assert not st.any()
    assert limbs_to_ints(out) == [opedersen(a, b) for a, b in zip(xs, ys)]
    # out of range in either position, and ragged validity inside one batch
    out, st = ctx.pedersen_hash2(ints_to_limbs([1, P, 2, 2**256 - 1]), ints_to_limbs([P, 1, 3, 1]))
    assert list(st) == [1, 1, 0, 1] and limbs_to_ints(out)[2] == opedersen(2, 3)
    assert limbs_to_ints(out)[0] == 0 and limbs_to_ints(out)[1] == 0       # failed elements are zeroed


def test_ntt_full_size_round_trip_and_dc_term(ctx):
    """2^22 points: inverse(forward) is the identity, and output 0 (frequency 0) is the sum of the inputs."""
    log_n = 22
    x = rand_felts(1 << log_n, 91)
    f = ctx.ntt(x, log_n, False, NTT_NAT_TO_REV)
    assert np.array_equal(ctx.ntt(f, log_n, True, NTT_REV_TO_NAT), x)
    assert limbs_to_ints(f[:1])[0] == sum(limbs_to_ints(x)) % P           # bit-reversed position 0 = frequency 0


def test_lde_interpolates_trace_on_coset_zero_shift(ctx):
    """LDE with offset 1: coset 0 reproduces the column itself (size-independent property, 2^16 x 3)."""
    log_n, C = 16, 3
    tr = rand_felts(C << log_n, 17)
    out = ctx.lde(tr, log_n, C, 3, offset=1).reshape(8, C, 1 << log_n, 4)
    assert np.array_equal(out[0].reshape(-1, 4), tr)
    assert np.array_equal(out.reshape(-1, 4), clib.lde(tr, log_n, C, 3, offset=1))


def root_dummy():
    import ctypes as C
    return np.empty(8, np.uint8).ctypes.data_as(C.c_void_p)


def test_argument_errors_are_reported_not_crashed(ctx):
    """Bad arguments come back as SPG_E_ARG with a message (the Python layer raises SpgError); the context stays usable."""
    import ctypes as C
    import stark_perpetual_b200 as spg
    lib, h = ctx._lib, ctx._h
    x = rand_felts(8, 1)
    assert lib.spg_ntt(h, x.ctypes.data_as(C.c_void_p), 27, 1, 0, 0, 0) == -2     # above the 2^26 twiddle table
    with pytest.raises(spg.SpgError, match="chain_len"):
        ctx.pedersen_chain(x, 0) if False else ctx._check(lib.spg_pedersen_chain_batch(h, x.ctypes.data_as(C.c_void_p), 0,
                                                          x.ctypes.data_as(C.c_void_p), root_dummy(), 1, 0))
    assert lib.spg_ntt(h, None, 3, 1, 0, 0, 0) == -2         # null data
    assert lib.spg_ntt(h, x.ctypes.data_as(C.c_void_p), 3, 1, 0, 7, 0) == -2      # bad order
    root = np.empty(32, np.uint8)
    assert lib.spg_merkle_commit(h, x.ctypes.data_as(C.c_void_p), 1, 6, root.ctypes.data_as(C.c_void_p), None, 0) == -2   # rows not 2^k
    ln = C.c_size_t(0)
    assert lib.spg_prove(h, x.ctypes.data_as(C.c_void_p), 8, 0, x.ctypes.data_as(C.c_void_p), 30, None, 0, C.byref(ln), 0) == -2   # log_n < 9
    st = np.zeros(1, np.uint8)
    assert lib.spg_pedersen_merkle_tree(h, x.ctypes.data_as(C.c_void_p), 6, root.ctypes.data_as(C.c_void_p), None,
                                        st.ctypes.data_as(C.c_void_p), 0) == -2                                         # 6 leaves
    assert b"bad argument" in lib.spg_last_error(h)
    vp = C.c_void_p
    fields = (vp * 11)(*([x.ctypes.data] * 11))
    assert lib.spg_message_hash_batch(h, 6, C.cast(fields, vp), x.ctypes.data_as(vp), st.ctypes.data_as(vp), 1, 0) == -2   # kind
    fields[1] = None
    assert lib.spg_message_hash_batch(h, 4, C.cast(fields, vp), x.ctypes.data_as(vp), st.ctypes.data_as(vp), 1, 0) == -2   # null array
    assert lib.spg_sign_batch(h, x.ctypes.data_as(vp), None, None, x.ctypes.data_as(vp), x.ctypes.data_as(vp),
                              st.ctypes.data_as(vp), 1, 0) == -2                                                        # null key
    # still healthy
    assert np.array_equal(ctx.ntt(ctx.ntt(x, 3, False, NTT_NAT_TO_REV), 3, True, NTT_REV_TO_NAT), x)
