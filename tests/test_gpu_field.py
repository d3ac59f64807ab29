"""Device field arithmetic (csrc/field.cuh) against the CPU oracle, through the C ABI's diagnostic entry point."""
import numpy as np
import pytest

import proof_systems_b200 as zk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def sample(orc, fid, n, seed):
    m = orc.MODULUS[fid]
    sp = [0, 1, 2, m - 1, m - 2, 1 << 254, (1 << 254) - 1, (1 << 32) - 1, 1 << 32, 1 << 224, m >> 1, (1 << 255) % m]
    a = orc.random_scalars(fid, n, seed)
    a[: len(sp)] = orc.ints_to_limbs(sp)
    return a


@pytest.mark.parametrize("fid", [0, 1])
def test_mul_add_sub_bit_exact(ctx, orc, fid):
    n = 4096
    a = sample(orc, fid, n, 1)
    b = sample(orc, fid, n, 2)[::-1].copy()
    for op, f in (("mul", orc.fe_mul), ("add", orc.fe_add), ("sub", orc.fe_sub)):
        got = ctx.field_op(fid, op, a, b)
        want = np.stack([f(fid, a[i], b[i]) for i in range(n)])
        assert np.array_equal(got, want), op


@pytest.mark.parametrize("fid", [0, 1])
def test_inverse(ctx, orc, fid):
    a = orc.to_mont(fid, sample(orc, fid, 64, 3))
    got = ctx.field_op(fid, "inv", a)
    want = np.stack([orc.fe_inv(fid, x) for x in a])
    assert np.array_equal(got, want)


def test_mul_throughput_reports_a_number(ctx):
    v = ctx.mul_throughput(zk.FP, 500)
    assert v > 1e9
    print(f"Fp Montgomery multiplications/s: {v:.3e}")
