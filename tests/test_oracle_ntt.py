"""CPU oracle, NTT layer (restates ark_poly::Radix2EvaluationDomain::{fft,ifft}_in_place).

The reference holds no numeric FFT vector; the conventions are pinned by (i) the definition-level DFT with
w = (5^T)^(2^(32-log n)), (ii) the Lagrange-basis golden vectors (tests/test_oracle_curve.py, which run the same
butterfly network over group elements), (iii) the domain chain of kimchi/src/circuits/domains.rs:40-69.
"""
import numpy as np
import pytest

FIELDS = [0, 1]


@pytest.mark.parametrize("fid", FIELDS)
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_ntt_matches_definition(orc, fid, log_n):
    n = 1 << log_n
    a = orc.to_mont(fid, orc.random_scalars(fid, n, seed=100 + log_n))
    assert np.array_equal(orc.ntt(fid, a), orc.dft_naive(fid, a))
    assert np.array_equal(orc.ntt(fid, a, inverse=True), orc.dft_naive(fid, a, inverse=True))


@pytest.mark.parametrize("fid", FIELDS)
@pytest.mark.parametrize("log_n", [1, 4, 11, 16])
def test_roundtrip_and_delta(orc, fid, log_n):
    n = 1 << log_n
    m = orc.MODULUS[fid]
    a = orc.to_mont(fid, orc.random_scalars(fid, n, seed=7))
    assert np.array_equal(orc.ntt(fid, orc.ntt(fid, a), inverse=True), a)
    assert np.array_equal(orc.ntt(fid, orc.ntt(fid, a, inverse=True)), a)
    # forward transform of delta_k is (w^{ik})_i
    k = 3 % n
    d = np.zeros((n, 4), dtype=np.uint64)
    d[k] = orc.fe(fid, 1)
    out = orc.ntt(fid, d)
    w = orc.fe_int(fid, orc.root_of_unity(fid, log_n))
    for i in sorted({0, 1, 2 % n, n // 2, n - 1}):
        assert orc.fe_int(fid, out[i]) == pow(w, i * k, m)


@pytest.mark.parametrize("fid", FIELDS)
def test_domain_chain_subsampling(orc, fid):
    """d1 c d2 c d4 c d8 (kimchi/src/circuits/domains.rs:64-66): evaluations over d8 sub-sampled by 8 are the
    evaluations over d1 (used at poly-commitment/src/ipa.rs:717-722)."""
    log_n = 6
    n = 1 << log_n
    coeffs = orc.to_mont(fid, orc.random_scalars(fid, n, seed=9))
    big = np.zeros((8 * n, 4), dtype=np.uint64)
    big[:n] = coeffs
    ev8 = orc.ntt(fid, big)
    ev1 = orc.ntt(fid, coeffs)
    assert np.array_equal(ev8[::8], ev1)
    m = orc.MODULUS[fid]
    g8 = orc.fe_int(fid, orc.root_of_unity(fid, log_n + 3))
    g1 = orc.fe_int(fid, orc.root_of_unity(fid, log_n))
    assert pow(g8, 8, m) == g1


@pytest.mark.parametrize("fid", FIELDS)
def test_coset_transform_definition(orc, fid):
    """coset FFT evaluates at 5*w^i (ark default coset offset = multiplicative generator, fp.rs:10)."""
    log_n = 4
    n = 1 << log_n
    m = orc.MODULUS[fid]
    c = orc.random_scalars(fid, n, seed=11)
    cm = orc.to_mont(fid, c)
    ev = orc.ntt(fid, cm, coset=True)
    w = orc.fe_int(fid, orc.root_of_unity(fid, log_n))
    ci = orc.limbs_to_ints(c)
    for i in [0, 1, 5, n - 1]:
        x = 5 * pow(w, i, m) % m
        assert orc.fe_int(fid, ev[i]) == sum(cj * pow(x, j, m) for j, cj in enumerate(ci)) % m
    assert np.array_equal(orc.ntt(fid, ev, inverse=True, coset=True), cm)
