"""CPU oracle, field layer: constants and arithmetic pinned against the reference's constants and Python ints.

Reference: curves/src/pasta/fields/fp.rs:8-80, fq.rs:8-79; curves/tests/pasta_curves.rs:10-33.
"""
import random

import numpy as np
import pytest

FIELDS = [0, 1]


@pytest.mark.parametrize("fid", FIELDS)
def test_montgomery_constants_rederived(orc, fid):
    m = orc.MODULUS[fid]
    assert orc.limbs_to_int(orc.const(fid, "modulus")) == m
    R = (1 << 256) % m
    assert orc.limbs_to_int(orc.const(fid, "one")) == R            # fp.rs:38-43 / fq.rs:36-41
    assert orc.limbs_to_int(orc.const(fid, "r2")) == (R * R) % m   # fp.rs:45-50 / fq.rs:43-48
    inv = (-pow(m, -1, 1 << 64)) % (1 << 64)
    assert inv == {0: 11037532056220336127, 1: 10108024940646105087}[fid]  # fp.rs:80 / fq.rs:79
    assert m.bit_length() == 255
    assert (m - 1) % (1 << 32) == 0 and ((m - 1) >> 32) % 2 == 1  # two-adicity 32


@pytest.mark.parametrize("fid", FIELDS)
def test_two_adic_root_is_5_pow_T(orc, fid):
    """TWO_ADIC_ROOT_OF_UNITY limbs (fp.rs:24-27 / fq.rs:21-24) == Montgomery form of 5^T, order exactly 2^32."""
    m = orc.MODULUS[fid]
    T = (m - 1) >> 32
    rho = pow(5, T, m)
    assert orc.fe_int(fid, orc.two_adic_root(fid)) == rho
    assert pow(rho, 1 << 32, m) == 1 and pow(rho, 1 << 31, m) != 1
    # SURVEY Appendix A derived values
    w16 = orc.fe_int(fid, orc.root_of_unity(fid, 16))
    assert w16 == pow(rho, 1 << 16, m)
    expect16 = {0: 0x23222d06029d21a655392ad9dda387278c1c46359289a4d3d465aafc06d1cf1a,
                1: 0x385e22fc1565ebd8a13142cc27b8876f05ec017404d761eff3a89df3fe315f99}[fid]
    assert w16 == expect16


@pytest.mark.parametrize("fid", FIELDS)
def test_field_ops_match_python_ints(orc, fid):
    m = orc.MODULUS[fid]
    rng = random.Random(1234 + fid)
    specials = [0, 1, 2, m - 1, m - 2, (m - 1) // 2, 1 << 254, (1 << 254) - 1]
    vals = specials + [rng.randrange(m) for _ in range(200)]
    for _ in range(400):
        a, b = rng.choice(vals), rng.choice(vals)
        A, B = orc.fe(fid, a), orc.fe(fid, b)
        assert orc.fe_int(fid, orc.fe_mul(fid, A, B)) == a * b % m
        assert orc.fe_int(fid, orc.fe_add(fid, A, B)) == (a + b) % m
        assert orc.fe_int(fid, orc.fe_sub(fid, A, B)) == (a - b) % m
    for a in vals:
        A = orc.fe(fid, a)
        got = orc.fe_int(fid, orc.fe_inv(fid, A))
        assert got == (pow(a, -1, m) if a else 0)
        s = orc.fe_sqrt(fid, orc.fe_mul(fid, A, A))
        assert s is not None and orc.fe_int(fid, s) in (a, (m - a) % m)
    # 5 generates the multiplicative group, hence is a non-residue
    assert orc.fe_sqrt(fid, orc.fe(fid, 5)) is None


def test_canonical_vs_montgomery_kat(orc):
    """curves/tests/pasta_curves.rs:10-33 — into_bigint returns canonical, storage is Montgomery."""
    y = 12418654782883325593414442427049395787963493412651469444558597405572177144507
    Y = orc.fe(orc.FP, y)
    assert orc.limbs_to_int(Y) == y * (1 << 256) % orc.FP_MODULUS  # raw limbs are Montgomery
    assert orc.fe_int(orc.FP, Y) == y
    assert orc.fe_int(orc.FP, orc.fe(orc.FP, 1)) == 1


@pytest.mark.parametrize("fid", FIELDS)
def test_vector_conversions_roundtrip(orc, fid):
    a = orc.random_scalars(fid, 1000, seed=5)
    assert np.array_equal(orc.from_mont(fid, orc.to_mont(fid, a)), a)
