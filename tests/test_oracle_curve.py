"""CPU oracle, group + MSM layer, pinned against every vector the reference holds for this path.

  curves/tests/pasta_curves.rs:35-74          Pallas affine-addition KAT
  curves/src/pasta/curves/{pallas,vesta}.rs   generators, cofactor 1
  kimchi/src/proof.rs:1163-1204               16-point Vesta MSM KAT
  srs/{pallas,vesta}.srs vs srs/test_*.srs    compressed vs uncompressed generators (point decompression)
  srs/test_{pallas,vesta}.srs                 lagrange_bases[n][i] == sum_j (w_n^{-ij}/n) g[j]: pinned n-point MSMs and
                                              the iFFT root/order/scale (poly-commitment/src/ipa.rs:1065-1172)
"""
import numpy as np
import pytest

PALLAS_GY = 12418654782883325593414442427049395787963493412651469444558597405572177144507
VESTA_GY = 11426906929455361843568202299992114520848200991084027513389447476559454104162


def gen(orc, cid):
    f = orc.BASE_FIELD[cid]
    return np.concatenate([orc.fe(f, 1), orc.fe(f, PALLAS_GY if cid == orc.PALLAS else VESTA_GY)])


def test_pallas_affine_add_kat(orc):
    f = orc.FP
    p1 = np.concatenate([orc.fe(f, 1), orc.fe(f, PALLAS_GY)])
    p2 = np.concatenate([
        orc.fe(f, 20444556541222657078399132219657928148671392403212669005631716460534733845831),
        orc.fe(f, PALLAS_GY)])
    assert orc.on_curve(orc.PALLAS, p1) and orc.on_curve(orc.PALLAS, p2)
    p3 = orc.affine_add(orc.PALLAS, p1, p2)
    assert orc.fe_int(f, p3[:4]) == 8503465768106391777493614032514048814691664078728891710322960303815233784505
    assert orc.fe_int(f, p3[4:]) == 16529367526445723262478303825122581175399563069290091271396079358777790485830


@pytest.mark.parametrize("cid", [0, 1])
def test_generator_and_group_order(orc, cid):
    G = gen(orc, cid)
    assert orc.on_curve(cid, G)
    r = orc.MODULUS[orc.SCALAR_FIELD[cid]]
    assert not np.any(orc.scalar_mul(cid, G, r))              # [r]G = identity (encoded as zeros)
    assert np.array_equal(orc.scalar_mul(cid, G, r + 1), G)
    assert np.array_equal(orc.scalar_mul(cid, G, 1), G)
    # doubling through the addition law's P == Q branch
    assert np.array_equal(orc.affine_add(cid, G, G), orc.scalar_mul(cid, G, 2))
    # P + (-P) = identity
    negG = G.copy()
    negG[4:] = orc.fe_sub(orc.BASE_FIELD[cid], np.zeros(4, dtype=np.uint64), G[4:])
    assert not np.any(orc.affine_add(cid, G, negG))
    # identity is neutral
    assert np.array_equal(orc.affine_add(cid, G, np.zeros(8, dtype=np.uint64)), G)
    assert np.array_equal(orc.affine_add(cid, np.zeros(8, dtype=np.uint64), G), G)


def b_poly_coefficients(chals, m):
    """poly-commitment/src/commitment.rs b_poly_coefficients: s_i = prod_{j: bit_j(i)=1} u_{k-j}"""
    k = len(chals)
    s = [1] * (1 << k)
    pw, kk = 1, 0
    for i in range(1, 1 << k):
        if i == pw * 2:
            pw *= 2
            kk += 1
        s[i] = s[i - pw] * chals[k - 1 - kk] % m
    return s


def test_b_poly_coefficients_kat():
    """poly-commitment/src/commitment.rs:869-910"""
    assert b_poly_coefficients([2, 3, 5, 7], 1 << 255) == [1, 7, 5, 35, 3, 21, 15, 105, 2, 14, 10, 70, 6, 42, 30, 210]


def test_vesta_msm_kat(orc):
    """kimchi/src/proof.rs:1163-1204 — basis i*G (i = 1..16), scalars b_poly_coefficients([2,3,5,7])."""
    cid = orc.VESTA
    G = gen(orc, cid)
    basis = np.stack([orc.scalar_mul(cid, G, i) for i in range(1, 17)])
    coeffs = b_poly_coefficients([2, 3, 5, 7], orc.FP_MODULUS)
    sc = orc.ints_to_limbs(coeffs)
    ex = 3756288960823668761746459900985719106126835112055076922409498125279524024429
    ey = 7540929664328976141648477194277016811781677917189411360504995258251130097840
    for algo in (0, 1):
        r = orc.msm(cid, basis, sc, algo=algo)
        assert orc.fe_int(orc.FQ, r[:4]) == ex and orc.fe_int(orc.FQ, r[4:]) == ey
    # VariableBaseMSM::msm takes Montgomery scalars
    r = orc.msm_mont(cid, basis, orc.to_mont(orc.FP, sc))
    assert orc.fe_int(orc.FQ, r[:4]) == ex


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_decompression_matches_uncompressed_srs(orc, request, name):
    srs = request.getfixturevalue(name)
    g = srs.g
    assert np.array_equal(g[:2048], srs.mont_points(srs.g_xy_canon))
    assert all(orc.on_curve(srs.cid, g[i]) for i in range(0, g.shape[0], 997))
    assert orc.on_curve(srs.cid, srs.mont_points(srs.h_xy_canon)[0])


def lagrange_scalars(orc, fid, n, i):
    """row i of the inverse DFT matrix: w_n^{-ij}/n, canonical"""
    m = orc.MODULUS[fid]
    log_n = n.bit_length() - 1
    w = orc.fe_int(fid, orc.root_of_unity(fid, log_n))
    wi = pow(w, -i, m) if n > 1 else 1
    ninv = pow(n, -1, m)
    out, cur = [], ninv
    for _ in range(n):
        out.append(cur)
        cur = cur * wi % m
    return orc.ints_to_limbs(out)


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_lagrange_small_domains_all_entries(orc, request, name):
    """n = 1..64: every entry, through the MSM (both algorithms) and through the group iFFT."""
    srs = request.getfixturevalue(name)
    for log_n in range(0, 7):
        n = 1 << log_n
        want = srs.lagrange_small(n)
        assert np.array_equal(orc.group_intt(srs.cid, srs.g[:n]), want), f"group iFFT n={n}"
        for i in range(n):
            sc = lagrange_scalars(orc, srs.scalar, n, i)
            assert np.array_equal(orc.msm(srs.cid, srs.g[:n], sc), want[i]), f"msm n={n} i={i}"
            if n <= 8:
                assert np.array_equal(orc.msm(srs.cid, srs.g[:n], sc, algo=1), want[i])


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_lagrange_2048_pinned_msms(orc, request, name):
    """BASELINE config 1: 2^11-point MSMs whose answers are stored in srs/test_*.srs."""
    srs = request.getfixturevalue(name)
    want = srs.mont_points(srs.lag_2048_canon)
    for i in [0, 1, 2, 1000, 2047]:
        sc = lagrange_scalars(orc, srs.scalar, 2048, i)
        assert np.array_equal(orc.msm(srs.cid, srs.g[:2048], sc), want[i])


def test_lagrange_2048_group_ifft_all_entries(orc, pallas_srs):
    """All 2048 entries at once through the oracle's butterfly network over group elements."""
    want = pallas_srs.mont_points(pallas_srs.lag_2048_canon)
    assert np.array_equal(orc.group_intt(orc.PALLAS, pallas_srs.g[:2048]), want)


def test_lagrange_65536_pinned_msms(orc, pallas_srs):
    """BASELINE config 2 size: 2^16-point Pallas MSMs with answers stored in srs/test_pallas.srs."""
    want = pallas_srs.mont_points(pallas_srs.lag_65536_canon)
    for k, i in enumerate(pallas_srs.lag_65536_idx[:3]):
        sc = lagrange_scalars(orc, orc.FQ, 65536, int(i))
        assert np.array_equal(orc.msm(orc.PALLAS, pallas_srs.g, sc), want[k])


@pytest.mark.parametrize("cid", [0, 1])
def test_pippenger_edge_cases_vs_definition(orc, cid, pallas_srs, vesta_srs):
    """SURVEY.md §8d edge set: scalars 0, 1, r-1, 2^k; repeated bases; P and -P; identity bases; n = 1; odd n."""
    srs = pallas_srs if cid == 0 else vesta_srs
    r = orc.MODULUS[srs.scalar]
    g = srs.g[:40].copy()
    g[5] = g[4]                      # repeated base
    g[7] = g[6]
    g[7, 4:] = orc.fe_sub(srs.base, np.zeros(4, dtype=np.uint64), g[6, 4:])   # -g[6]
    g[9] = 0                         # identity base (poly-commitment/src/ipa.rs:848-850 pads with zero())
    sc = orc.random_scalars(srs.scalar, 40, seed=3)
    sc[0] = 0
    sc[1] = orc.int_to_limbs(1)
    sc[2] = orc.int_to_limbs(r - 1)
    sc[3] = orc.int_to_limbs(1 << 200)
    sc[4] = sc[5] = orc.int_to_limbs(12345)
    sc[6] = sc[7] = orc.int_to_limbs(999)       # cancels
    for n in [1, 2, 3, 7, 33, 40]:
        assert np.array_equal(orc.msm(cid, g[:n], sc[:n]), orc.msm(cid, g[:n], sc[:n], algo=1)), n
    # all-zero scalars and empty input -> identity
    assert not np.any(orc.msm(cid, g[:8], np.zeros((8, 4), dtype=np.uint64)))
    assert not np.any(orc.msm(cid, g[:0], sc[:0]))
    # the reference's 2-way split (ipa.rs:652-662) gives the same group element
    assert np.array_equal(orc.msm_split2(cid, srs.g[:2048], orc.random_scalars(srs.scalar, 2048, seed=4), threads=4),
                          orc.msm(cid, srs.g[:2048], orc.random_scalars(srs.scalar, 2048, seed=4)))
    # thread count does not change the result
    big_sc = orc.random_scalars(srs.scalar, 2048, seed=4)
    assert np.array_equal(orc.msm(cid, srs.g[:2048], big_sc, threads=1), orc.msm(cid, srs.g[:2048], big_sc, threads=4))


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_extend_bases_are_curve_points_and_deterministic(orc, request, name):
    srs = request.getfixturevalue(name)
    ext = orc.extend_bases(srs.cid, srs.g[:64], 300)
    assert np.array_equal(ext[:64], srs.g[:64])
    assert all(orc.on_curve(srs.cid, p) for p in ext)
    for k, i in [(1, 0), (2, 5), (4, 43)]:
        assert np.array_equal(ext[k * 64 + i], orc.affine_add(srs.cid, ext[(k - 1) * 64 + i], srs.g[(i + k) % 64]))
    assert len({p.tobytes() for p in ext}) == 300          # no repeats
    assert np.array_equal(orc.extend_bases(srs.cid, srs.g[:64], 300), ext)
