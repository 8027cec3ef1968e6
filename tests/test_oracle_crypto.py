"""CPU: the Python oracle against every golden vector generated from the reference
(tests/golden/gen_golden.py) and the reference's own KATs."""
import pytest

from oracle import ecdsa, params, pedersen


def test_constants_structure():
    # signature.py:55-68
    assert 2**251 < params.EC_ORDER < params.FIELD_PRIME
    assert len(params.CONSTANT_POINTS) == 506
    assert params.SHIFT_POINT == (
        0x49EE3EBA8C1600700EE1B87EB599F16716B0B1022947733551FDE4050CA6804,
        0x3CA0CFE4B3BC6DDF346D49D06EA0ED34E621062C0E056C1D0405D266E10268A)
    assert params.EC_GEN == (
        0x1EF15C18599971B7BECED415A40F0C7DEACFD9B0D1819E03D723D8BC943CFCA,
        0x5668060AA49730B7BE4801DF46EC62DE53ECD11ABE43A32873000C36E8DC1F)
    for x, y in params.CONSTANT_POINTS:
        assert ecdsa.is_point_on_curve(x, y)
    assert params.R_MOD_P == 0x7fffffffffffdf0ffffffffffffffffffffffffffffffffffffffffffffffe1
    assert pow(params.root_of_unity(18), 1 << 17, params.FIELD_PRIME) == params.FIELD_PRIME - 1


def test_pedersen_golden(golden):
    for a, b, o, _tag in golden["pedersen"]:
        assert pedersen.pedersen_hash(int(a, 16), int(b, 16)) == int(o, 16)
    for a, o in golden["pedersen_single"]:
        assert pedersen.pedersen_hash(int(a, 16)) == int(o, 16)


def test_pedersen_range_and_bytes():
    with pytest.raises(AssertionError):
        pedersen.pedersen_hash(params.FIELD_PRIME, 1)
    with pytest.raises(AssertionError):
        pedersen.pedersen_hash(1, -1)
    x = (5).to_bytes(32, "big")
    assert pedersen.pedersen_hash_func(x, x) == pedersen.pedersen_hash(5, 5).to_bytes(32, "big")


def test_keys_golden(golden):
    for priv, pub in golden["keys"][:12]:
        assert ecdsa.private_to_stark_key(int(priv, 16)) == int(pub, 16)


def test_verify_golden(golden):
    for msg, r, s, pub, res, tag in golden["verify"]:
        pk = int(pub, 16) if isinstance(pub, str) else (int(pub[0], 16), int(pub[1], 16))
        try:
            got = 1 if ecdsa.verify(int(msg, 16), int(r, 16), int(s, 16), pk) else 0
        except AssertionError:
            got = 2
        assert got == res, tag


def test_sign_golden(golden):
    for mh, priv, er, es in golden["sign_js_kat"]:
        assert ecdsa.sign(int(mh, 16), int(priv, 16)) == (int(er, 16), int(es, 16))
    for msg, priv, r, s in golden["sign"][:6]:
        assert ecdsa.sign(int(msg, 16), int(priv, 16)) == (int(r, 16), int(s, 16))


def test_get_y_and_grind(golden):
    for x, y in golden["get_y"]:
        if y is None:
            with pytest.raises(ecdsa.InvalidPublicKeyError):
                ecdsa.get_y_coordinate(int(x, 16))
            assert not ecdsa.is_valid_stark_key(int(x, 16))
        else:
            assert ecdsa.get_y_coordinate(int(x, 16)) == int(y, 16)
    for a, b, o in golden["grind_key"]:
        assert ecdsa.grind_key(int(a, 16), int(b, 16)) == int(o, 16)
