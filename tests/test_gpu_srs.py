"""The SRS mirror (proof_systems_b200.SRS over srs.cu) against the reference's contracts:
  poly-commitment/src/pbt_srs.rs:21-85       chunk-count contract of commit_non_hiding
  poly-commitment/tests/ipa_commitment.rs:27-119   interpolate+commit == commit_evaluations on the Lagrange basis
  poly-commitment/src/ipa.rs:605-622,643-647       mask_custom, zero polynomial
"""
import numpy as np
import pytest

import proof_systems_b200 as zk
from proof_systems_b200.host import BlindersDontMatch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_commit_non_hiding_expected_number_of_chunks(ctx, orc, request, name):
    g = request.getfixturevalue(name)
    rng = np.random.default_rng(7)
    for log2_srs_size in (1, 3, 5):
        n = 1 << log2_srs_size
        srs = zk.SRS(ctx, g.cid, g.g[:n], g.mont_points(g.h_xy_canon)[0])
        assert srs.max_poly_size() == n

        def rand_poly(k):
            return orc.to_mont(g.scalar, orc.random_scalars(g.scalar, k, seed=int(rng.integers(1 << 30))))

        assert len(srs.commit_non_hiding(rand_poly(n), 1)) == 1
        k = int(rng.integers(2, 10))
        assert len(srs.commit_non_hiding(rand_poly(n), k)) == k
        k = int(rng.integers(1, 10))
        zero = srs.commit_non_hiding(np.zeros((0, 4), dtype=np.uint64), k)
        assert len(zero) == k and not np.any(zero.chunks)                    # ipa.rs:643-647
        k = int(rng.integers(2, 5))
        p = rand_poly(k * n)
        assert len(srs.commit_non_hiding(p, int(rng.integers(1, k)))) == k
        assert len(srs.commit_non_hiding(p, k)) == k
        req = int(rng.integers(k + 1, 10))
        c = srs.commit_non_hiding(p, req)
        assert len(c) == req and not np.any(c.chunks[k:])
        # values: chunk j is the MSM of coefficient block j (incl. a ragged last block)
        p2 = rand_poly(2 * n + max(1, n // 2))
        c2 = srs.commit_non_hiding(p2, 1)
        assert len(c2) == 3
        for j in range(3):
            blk = orc.from_mont(g.scalar, p2[j * n:(j + 1) * n])
            assert np.array_equal(c2.chunks[j], orc.msm(g.cid, g.g[:len(blk)], blk)), j
        # an UNTRIMMED coefficient vector (zero tail chunk): the chunk count follows plnm.len() as given (ipa.rs:646-676),
        # the tail chunk is the identity; an all-zero vector is is_zero() -> one identity chunk
        p3 = rand_poly(3 * n)
        p3[2 * n:] = 0
        c3 = srs.commit_non_hiding(p3, 1)
        assert len(c3) == 3 and not np.any(c3.chunks[2]) and np.array_equal(c3.chunks[:2], srs.commit_non_hiding(p3[:2 * n], 1).chunks)
        z3 = srs.commit_non_hiding(np.zeros((3 * n, 4), dtype=np.uint64), 1)
        assert len(z3) == 1 and not np.any(z3.chunks)
        srs.close()


def test_commit_evaluations_equals_interpolate_then_commit(ctx, orc, pallas_srs):
    """ipa_commitment.rs:27-119 (single-chunk case): <evals, lagrange_basis> == commit(interpolate(evals))."""
    g = pallas_srs
    n = 1 << 10
    srs = zk.SRS(ctx, g.cid, g.g[:n], g.mont_points(g.h_xy_canon)[0])
    srs.add_lagrange_basis(n, g.lagrange_small(n))          # the reference's stored basis for n = 1024
    evals = orc.to_mont(g.scalar, orc.random_scalars(g.scalar, n, seed=21))
    dom = zk.Radix2EvaluationDomain(ctx, g.scalar, n)
    coeffs = dom.ifft(evals)
    via_coeffs = srs.commit_non_hiding(coeffs, 1)
    via_evals = srs.commit_evaluations_non_hiding(n, evals)
    assert np.array_equal(via_coeffs.chunks, via_evals.chunks)
    # evaluations living on a 4x larger domain are sub-sampled (ipa.rs:717-722)
    big = np.zeros((4 * n, 4), dtype=np.uint64)
    big[:n] = coeffs
    evals4 = zk.Radix2EvaluationDomain(ctx, g.scalar, 4 * n).fft(big)
    assert np.array_equal(srs.commit_evaluations_non_hiding(n, evals4).chunks, via_evals.chunks)
    # the reference panics when the commitment domain is larger than the evaluation domain
    with pytest.raises(zk.ZkError):
        srs.commit_evaluations_non_hiding(n, evals[: n // 2])
    srs.close()


def test_mask_custom(ctx, orc, vesta_srs):
    g = vesta_srs
    n = 64
    h = g.mont_points(g.h_xy_canon)[0]
    srs = zk.SRS(ctx, g.cid, g.g[:n], h)
    p = orc.to_mont(g.scalar, orc.random_scalars(g.scalar, 2 * n, seed=4))
    com = srs.commit_non_hiding(p, 2)
    bl_can = orc.random_scalars(g.scalar, 2, seed=5)
    blinders = orc.to_mont(g.scalar, bl_can)
    masked = srs.mask_custom(com, blinders)
    for j in range(2):
        want = orc.affine_add(g.cid, com.chunks[j], orc.scalar_mul(g.cid, h, orc.limbs_to_int(bl_can[j])))
        assert np.array_equal(masked.chunks[j], want)
    with pytest.raises(BlindersDontMatch):
        srs.mask_custom(com, blinders[:1])
    srs.close()


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_device_point_decompression(ctx, orc, request, name):
    """srs/{pallas,vesta}.srs store compressed generators (utils/src/serialization.rs:65-106); the device decoder must
    reproduce the uncompressed coordinates of srs/test_*.srs and agree with the oracle on every point, incl. infinity."""
    g = request.getfixturevalue(name)
    got = ctx.decompress_points(g.cid, g.g_cmp)
    assert np.array_equal(got[:2048], g.mont_points(g.g_xy_canon))
    assert np.array_equal(got, g.g)                       # oracle decompression of the whole fixture
    inf = bytes(32) + bytes([0x40])
    mixed = g.g_cmp[:3].tobytes() + inf + g.g_cmp[3:5].tobytes()
    out = ctx.decompress_points(g.cid, mixed)
    assert not np.any(out[3]) and np.array_equal(out[4], g.g[3])
    # flipping the sign flag gives the negated point; an x that is off the curve is rejected
    flipped = bytearray(g.g_cmp[0].tobytes()); flipped[32] ^= 0x80
    neg = ctx.decompress_points(g.cid, bytes(flipped))[0]
    assert np.array_equal(neg[:4], g.g[0][:4]) and not np.any(orc.affine_add(g.cid, neg, g.g[0]))
    bad = None
    for x in range(2, 40):   # find a small x with x^3 + 5 a non-residue
        if orc.fe_sqrt(g.base, orc.fe_add(g.base, orc.fe_mul(g.base, orc.fe_mul(g.base, orc.fe(g.base, x), orc.fe(g.base, x)), orc.fe(g.base, x)), orc.fe(g.base, 5))) is None:
            bad = x.to_bytes(32, "little") + bytes([0])
            break
    assert bad is not None
    with pytest.raises(zk.ZkError):
        ctx.decompress_points(g.cid, bad)
    # non-canonical encodings are rejected like ark-serialize does (SWFlags::from_u8 + the field's canonical check): x = p,
    # x = p + 1 (both reduce to valid small x values: 0 is off-curve, but p + 1 == 1 is the generator's x), stray low flag
    # bits, and infinity together with the sign bit
    p_int = orc.FP_MODULUS if g.base == orc.FP else orc.FQ_MODULUS
    good = g.g_cmp[0].tobytes()
    cases = [p_int.to_bytes(32, "little") + bytes([0]), (p_int + 1).to_bytes(32, "little") + bytes([0]),
             good[:32] + bytes([good[32] | 0x01]), good[:32] + bytes([good[32] | 0x20]), bytes(32) + bytes([0xc0])]
    for enc in cases:
        with pytest.raises(zk.ZkError):
            ctx.decompress_points(g.cid, good + enc)


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_point_codecs_round_trip_the_reference_bytes(ctx, orc, request, name):
    """compress(decompress(raw)) == raw on every generator the reference ships (srs/*.srs bytes), and the uncompressed form of
    srs/test_*.srs decodes to the same points (utils/src/serialization.rs:65-146)."""
    G = request.getfixturevalue(name)
    pts = ctx.decompress_points(G.cid, G.g_cmp)
    assert np.array_equal(ctx.compress_points(G.cid, pts), G.g_cmp)
    raw65 = np.concatenate([G.g_xy_canon, np.zeros((2048, 1), dtype=np.uint8)], axis=1)
    assert np.array_equal(ctx.points_from_uncompressed(G.cid, raw65), pts[:2048])
    # the identity: flag bit 6
    ident = np.zeros((1, 33), dtype=np.uint8)
    ident[0, 32] = 0x40
    assert not np.any(ctx.decompress_points(G.cid, ident))
    assert np.array_equal(ctx.compress_points(G.cid, np.zeros((1, 8), dtype=np.uint64)), ident)
    raw65[5, 64] = 0x40
    assert not np.any(ctx.points_from_uncompressed(G.cid, raw65[:8])[5])
    # a coordinate >= the modulus is rejected
    raw65[6, :32] = 0xFF
    with pytest.raises(zk.ZkError):
        ctx.points_from_uncompressed(G.cid, raw65[:8])


def test_srs_from_file(ctx, orc, vesta_srs, tmp_path):
    """get_srs_test (precomputed_srs.rs:84-91): g, h and the stored Lagrange bases straight from a test_*.srs-format file;
    commit_evaluations then uses the file's basis, which equals the one built on the device."""
    from proof_systems_b200 import srs_file
    G = vesta_srs
    n = 1024
    flag = np.zeros((n, 1), dtype=np.uint8)
    g65 = np.concatenate([G.g_xy_canon[:n], flag], axis=1)
    h65 = np.concatenate([G.h_xy_canon, [0]]).astype(np.uint8)
    bases = {m: np.concatenate([G.lag_small_canon[m - 1:2 * m - 1], flag[:m]], axis=1).reshape(m, 1, 65) for m in (1, 4, n)}
    path = str(tmp_path / "test_vesta.srs")
    srs_file.write_srs(path, srs_file.SrsFile(g=g65, h=h65, lagrange_bases=bases))
    srs = zk.SRS.from_file(ctx, G.cid, path)
    assert srs.max_poly_size() == n and np.array_equal(srs.g, G.g[:n]) and np.array_equal(srs.h, G.mont_points(G.h_xy_canon)[0])
    ev = orc.to_mont(G.scalar, orc.random_scalars(G.scalar, n, seed=77))
    got = srs.commit_evaluations_non_hiding(n, ev)
    want = orc.msm_mont(G.cid, G.lagrange_small(n), ev)
    assert np.array_equal(got.chunks[0], want)
    assert np.array_equal(srs.get_lagrange_basis_from_domain_size(n), G.lagrange_small(n))   # served from the file's copy
    fresh = zk.SRS.from_file(ctx, G.cid, path, lagrange=False)
    assert np.array_equal(fresh.get_lagrange_basis_from_domain_size(n), G.lagrange_small(n))  # built on the device
    srs.close()
    fresh.close()


def test_commit_evaluations_batch_and_lanes(ctx, orc, pallas_srs):
    """15 witness-like columns in one call (kimchi/src/prover.rs:329-351): identical to 15 single calls, whatever number of MSMs
    is fused into one pipeline (16 = all of them, 1 = one pipeline per MSM, 4 = ragged groups)."""
    G = pallas_srs
    n = 2048
    srs = zk.SRS(ctx, G.cid, G.g[:n], G.mont_points(G.h_xy_canon)[0])
    srs.add_lagrange_basis(n, G.mont_points(G.lag_2048_canon))
    ev = orc.to_mont(G.scalar, orc.random_scalars(G.scalar, 15 * n, seed=5)).reshape(15, n, 4)
    ev[3] = 0
    ev[4, : n - 5] = orc.to_mont(G.scalar, orc.ints_to_limbs([1]))[0]
    want = [orc.msm_mont(G.cid, G.mont_points(G.lag_2048_canon), ev[j]) for j in range(15)]
    try:
        for lanes in (16, 1, 4):
            ctx.set_option("msm_batch", lanes)
            got = srs.commit_evaluations_non_hiding_batch(n, ev)
            for j in range(15):
                assert np.array_equal(got[j].chunks[0], want[j]), (lanes, j)
        single = srs.commit_evaluations_non_hiding(n, ev[7])
        assert np.array_equal(single.chunks[0], want[7])
        # chunked commit_non_hiding (7 chunks of t) goes through the same fused pipeline
        c7 = srs.commit_non_hiding(ev[:7].reshape(-1, 4), 7)
        for j in range(7):
            assert np.array_equal(c7.chunks[j], orc.msm_mont(G.cid, G.g[:n], ev[j]))
        with pytest.raises(zk.ZkError):
            ctx.set_option("msm_batch", 17)
    finally:
        ctx.set_option("msm_batch", 16)
    srs.close()


def test_one_generator_srs_regression_bytes_on_device(ctx, orc):
    """precomputed_srs.rs:139-155: the serialised one-generator SRS decodes to the curve generator (pallas.rs:10-15,
    vesta.rs:10-15) and re-encodes to the same bytes, on both curves."""
    from test_srs_file import GENERATOR_Y, SRS_ONE_GENERATOR_HEX
    raw = np.frombuffer(bytes.fromhex(SRS_ONE_GENERATOR_HEX)[4:37], dtype=np.uint8).reshape(1, 33)
    for name, cid, fid in (("pallas", zk.PALLAS, orc.FP), ("vesta", zk.VESTA, orc.FQ)):
        pt = ctx.decompress_points(cid, raw)
        assert orc.limbs_to_ints(orc.from_mont(fid, pt.reshape(2, 4))) == [1, GENERATOR_Y[name]]
        assert np.array_equal(ctx.compress_points(cid, pt), raw)


def test_polycomm_serialization_regression_bytes_on_device(ctx, orc, vesta_srs):
    """ser_regression_canonical_polycomm (poly-commitment/tests/commitment.rs:345-385): srs.commit(DensePolynomial::rand(300, rng),
    6, rng) on SRS::<Vesta>::create(128) with StdRng seed [0; 32] — commit_non_hiding, mask_custom and the point compression all
    on the device path — serialises to the reference's hard-coded bytes."""
    import json
    import os

    from test_ser_regression import GOLDEN, msgpack_points, padded, polycomm_inputs
    G = vesta_srs
    coeffs, blinders = polycomm_inputs(orc)
    srs = zk.SRS(ctx, G.cid, G.g[:128], G.mont_points(G.h_xy_canon)[0])
    com = srs.commit_custom(coeffs, 6, blinders)                         # == srs.commit(&poly, 6, rng) with those blinders
    assert len(com) == 6
    pts33 = ctx.compress_points(G.cid, com.chunks)
    raw = msgpack_points([bytes(p) for p in pts33], struct_prefix=b"\x91")
    want = json.load(open(GOLDEN))["polycomm_vesta_srs128_deg300_chunks6"]
    assert padded(raw, len(want)) == want
    with pytest.raises(BlindersDontMatch):
        srs.commit_custom(coeffs, 6, blinders[:5])
    srs.close()


def test_opening_proof_serialization_regression_bytes_on_device(ctx, orc, vesta_srs):
    """ser_regression_canonical_opening_proof (poly-commitment/tests/commitment.rs:388-443): SRS::open with the seven folding
    rounds — the L/R MSMs, the inner products, the folds of a and b, the final sg — run by the device (zk.IpaRounds, csrc/ipa.cu)
    and the host-side transcript replayed around them (tests/open_replay.py) serialises to the reference's hard-coded bytes."""
    import json

    from open_replay import DeviceRounds, first_opening_proof_bytes
    from test_ser_regression import GOLDEN, padded
    raw = first_opening_proof_bytes(orc, vesta_srs, lambda g, a, b: DeviceRounds(orc, zk, ctx, g, a, b))
    want = json.load(open(GOLDEN))["opening_proof_vesta_srs128"]
    assert padded(raw, len(want)) == want


def test_opening_proof_regression_bytes_through_zk_srs_open(ctx, orc, vesta_srs):
    """The same regression (commitment.rs:388-443) through the PRODUCT-LEVEL entry point zk_srs_open (csrc/open.cu): combine_polys,
    b_init, the combined inner product, the rounds with h and U inside the MSMs, r_prime, delta, z1 and z2 are all computed by the
    library; the test only plays the transcript (sponge, group map) behind the callbacks and serialises the result."""
    import json

    from open_replay import opening_proof_bytes_product_level
    from test_ser_regression import GOLDEN, padded
    raw = opening_proof_bytes_product_level(orc, zk, ctx, vesta_srs)
    want = json.load(open(GOLDEN))["opening_proof_vesta_srs128"]
    assert padded(raw, len(want)) == want


@pytest.mark.parametrize("window_bits", [-1, 0])
def test_open_evaluation_form_and_chunked_inputs(ctx, orc, pallas_srs, window_bits):
    """combine_polys (poly-commitment/src/utils.rs:103-202) on the device: evaluation-form entries (sub-sampled, interpolated, chunked
    and linearised with powers of polyscale starting at 1, utils.rs:183-199) must give the proof of the equivalent
    coefficient-form batch; chunked coefficient-form entries and a non power-of-two SRS (padding, ipa.rs:848-862) are in the mix.
    The transcript behind the callbacks is a deterministic stand-in: equality of two library runs is what is checked, plus
    sg = <b_poly_coefficients(chals), g> and the verifier's z1 / z2 identity against the oracle."""
    G = pallas_srs
    fs, srs_len, dom = G.scalar, 48, 128
    m = orc.MODULUS[fs]
    ints = lambda a: orc.limbs_to_ints(orc.from_mont(fs, np.ascontiguousarray(a).reshape(-1, 4)))
    mont = lambda xs: orc.to_mont(fs, orc.ints_to_limbs([x % m for x in xs])) if len(xs) else np.zeros((0, 4), dtype=np.uint64)
    rnd = lambda k, seed: orc.to_mont(fs, orc.random_scalars(fs, k, seed=seed))
    srs = zk.SRS(ctx, G.cid, G.g[:srs_len], G.mont_points(G.h_xy_canon)[0], window_bits=window_bits)
    rounds = 6                                                            # ceil_log2(48)
    polyscale, evalscale = rnd(1, 1)[0], rnd(1, 2)[0]
    ps = ints(polyscale)[0]
    elm, draws = rnd(3, 3), rnd(2 * rounds + 2, 4)
    ev = rnd(4 * dom, 5)                                                  # evaluations on a 4x larger domain: stride 4
    ev_bl, dense, dense_bl = rnd(3, 6), rnd(100, 7), rnd(3, 8)            # domain 128 over |g| = 48: 3 chunks each
    chals = []

    def transcript():
        chals.clear()
        state = [7]

        def nxt():
            state[0] = (state[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            return state[0]
        u_base = lambda cip: G.g[100 + int(ints(cip)[0] % 50)]           # any curve point
        def rc(i, l, r):
            u = mont([nxt() * (1 << 64) + nxt() + ints(l[:4])[0] % 3])[0]
            chals.append(ints(u)[0])
            return u
        fc = lambda delta: mont([nxt() + 1])[0]
        return u_base, rc, fc

    proof_e = zk.srs_open(srs, [(dense, 0, dense_bl), (ev, dom, ev_bl)], elm, polyscale, evalscale, draws, *transcript())
    # the equivalent batch in coefficient form: the interpolated, chunk-linearised polynomial in the evaluation entry's place,
    # followed by empty entries that advance the scale like its remaining chunks (utils.rs:151-164)
    coeffs = ints(orc.ntt(fs, np.ascontiguousarray(ev[::4]), inverse=True))
    lin = [sum(pow(ps, k, m) * (coeffs[k * srs_len + i] if k * srs_len + i < dom else 0) for k in range(3)) % m for i in range(srs_len)]
    empty = np.zeros((0, 4), dtype=np.uint64)
    proof_c = zk.srs_open(srs, [(dense, 0, dense_bl), (mont(lin), 0, ev_bl[:1]), (empty, 0, ev_bl[1:2]), (empty, 0, ev_bl[2:3])], elm, polyscale,
                          evalscale, draws, *transcript())
    for a, b in ((proof_e.lr, proof_c.lr), (proof_e.delta, proof_c.delta), (proof_e.z1, proof_c.z1), (proof_e.z2, proof_c.z2), (proof_e.sg, proof_c.sg)):
        assert np.array_equal(a, b)
    assert proof_e.lr.shape == (rounds, 2, 8) and np.any(proof_e.lr[-1])
    # sg = <b_poly_coefficients(chals), g> (commitment.rs:565-581), an MSM the oracle can redo
    s = [1]
    for u in chals:
        s = [v for t in s for v in (t, t * u % m)]
    assert np.array_equal(proof_e.sg, orc.msm(G.cid, G.g[:srs_len], orc.ints_to_limbs(s[:srs_len])))
    srs.close()
