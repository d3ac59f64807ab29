"""SRS::lagrange_basis on the device (group inverse FFT of the generators, poly-commitment/src/ipa.rs:1065-1172) against
the reference's own stored bases in srs/test_{pallas,vesta}.srs (tests/golden) — the file the reference's
heavy_test_srs_serialization (precomputed_srs.rs:156-234) regenerates and compares."""
import numpy as np
import pytest

import proof_systems_b200 as zk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_lagrange_bases_match_reference_files(ctx, orc, request, name):
    g = request.getfixturevalue(name)
    srs = zk.SRS(ctx, g.cid, g.g[:2048], g.mont_points(g.h_xy_canon)[0])
    for log_n in range(0, 11):
        n = 1 << log_n
        assert np.array_equal(srs.get_lagrange_basis_from_domain_size(n), g.lagrange_small(n)), n
    assert np.array_equal(srs.get_lagrange_basis_from_domain_size(2048), g.mont_points(g.lag_2048_canon))
    # the registered basis is what commit_evaluations_non_hiding uses: <evals, L> == commit(interpolate(evals))
    evals = orc.to_mont(g.scalar, orc.random_scalars(g.scalar, 2048, seed=31))
    coeffs = zk.Radix2EvaluationDomain(ctx, g.scalar, 2048).ifft(evals)
    assert np.array_equal(srs.commit_evaluations_non_hiding(2048, evals).chunks, srs.commit_non_hiding(coeffs, 1).chunks)
    # domains larger than the SRS get chunked bases (ipa.rs:1145-1171): two chunks of 4096 points here
    assert srs.lagrange_basis_chunks(4096) == 2 and srs.get_lagrange_basis_from_domain_size(4096).shape == (2, 4096, 8)
    srs.close()


def test_lagrange_basis_2_16_sampled(ctx, orc, pallas_srs):
    """the prover's domain size: 65 536 commitments, each a 2^16-point MSM on the CPU side"""
    g = pallas_srs
    srs = zk.SRS(ctx, g.cid, g.g, g.mont_points(g.h_xy_canon)[0], window_bits=0)
    basis = srs.get_lagrange_basis_from_domain_size(1 << 16)
    want = g.mont_points(g.lag_65536_canon)
    for k, i in enumerate(g.lag_65536_idx):
        assert np.array_equal(basis[int(i)], want[k]), int(i)
    assert all(orc.on_curve(g.cid, basis[i]) for i in range(0, 1 << 16, 4099))
    srs.close()


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_chunked_lagrange_basis_and_commitment(ctx, orc, request, name):
    """A domain larger than the SRS (poly-commitment/src/ipa.rs:1145-1171): chunk c of element i is
    sum_j (w^{-i (c |g| + j)} / n) g[j] over the chunk's terms — every entry checked against that MSM on the oracle — and
    commit_evaluations_non_hiding returns one point per chunk, each the MSM of the evaluations against that chunk of the basis
    (PolyComm::multi_scalar_mul, commitment.rs:350-394), equal to committing the interpolated polynomial chunk by chunk."""
    G = request.getfixturevalue(name)
    srs_len, n = 24, 64                                  # 3 chunks: 24 + 24 + 16 terms
    srs = zk.SRS(ctx, G.cid, G.g[:srs_len], G.mont_points(G.h_xy_canon)[0])
    assert srs.lagrange_basis_chunks(n) == 3 and srs.lagrange_basis_chunks(16) == 1
    basis = srs.get_lagrange_basis_from_domain_size(n)
    assert basis.shape == (3, n, 8)
    m = orc.MODULUS[G.scalar]
    w = orc.fe_int(G.scalar, orc.root_of_unity(G.scalar, 6))
    winv, ninv = pow(w, -1, m), pow(n, -1, m)
    for c in range(3):
        terms = min(srs_len, n - c * srs_len)
        for i in (0, 1, 17, 63):
            scal = [pow(winv, i * (c * srs_len + j), m) * ninv % m for j in range(terms)]
            assert np.array_equal(basis[c, i], orc.msm(G.cid, G.g[:terms], orc.ints_to_limbs(scal))), (c, i)
    evals = orc.to_mont(G.scalar, orc.random_scalars(G.scalar, n, seed=44))
    com = srs.commit_evaluations_non_hiding(n, evals)
    assert len(com) == 3
    coeffs = orc.ntt(G.scalar, evals, inverse=True)
    via_coeffs = srs.commit_non_hiding(coeffs, 1)        # 64 coefficients over |g| = 24: the same three chunks
    assert np.array_equal(com.chunks, via_coeffs.chunks)
    srs.close()
