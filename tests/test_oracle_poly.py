"""The oracle's divide_by_vanishing (oracle/ntt_impl.h, restating ark-poly's DensePolynomial::divide_by_vanishing_poly as called at
kimchi/src/prover.rs:909) pinned by the identity it must satisfy, f = q (x^n - 1) + r with deg r < n, checked with Python integers."""
import pytest


@pytest.mark.parametrize("log_n,length", [(3, 8 * 7 + 3), (3, 64), (3, 8), (3, 5), (2, 32), (4, 33), (0, 9)])
def test_division_identity(orc, log_n, length):
    fid, P, n = orc.FQ, orc.FQ_MODULUS, 1 << log_n
    f = orc.to_mont(fid, orc.random_scalars(fid, length, seed=length))
    q, r = orc.divide_by_vanishing(fid, f, log_n)
    assert q.shape[0] == max(length - n, 0) and r.shape[0] == n
    ints = lambda a: orc.limbs_to_ints(orc.from_mont(fid, a)) if len(a) else []
    back = [0] * max(length, n)
    for j, c in enumerate(ints(q)):
        back[j + n] = (back[j + n] + c) % P
        back[j] = (back[j] - c) % P
    for j, c in enumerate(ints(r)):
        back[j] = (back[j] + c) % P
    assert back[:length] == ints(f) and not any(back[length:])
