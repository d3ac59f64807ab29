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
    # domains larger than the SRS need chunked bases: not on the device path
    with pytest.raises(zk.ZkError):
        srs.get_lagrange_basis_from_domain_size(4096)
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
