"""The reference's byte-exact serialization regressions on the commit path, reproduced with the oracle
(poly-commitment/tests/commitment.rs:288-385; SURVEY.md §8c vector 6):
  ser_regression_canonical_srs       SRS::create_trusted_setup_with_toxic_waste(Fp::rand(rng), 8): g[i] = x^i * G, h (ipa.rs:515-545)
  ser_regression_canonical_polycomm  srs.commit(DensePolynomial::rand(300, rng), 6, rng) on SRS::<Vesta>::create(128):
                                     chunking of commit_non_hiding (ipa.rs:638-683) + mask (ipa.rs:605-622, 686-693)
both driven by StdRng seed [0; 32] (tests/rust_rng.py restates rand's ChaCha12 stream and ark-ff's field sampling).  The bytes
are the reference's own (tests/golden/ser_regression.json, extracted by tests/golden/make_ser_regression.py)."""
import json
import os

import numpy as np
import pytest

from rust_rng import StdRng

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ser_regression.json")
GENERATOR_Y = {"pallas": 12418654782883325593414442427049395787963493412651469444558597405572177144507,
               "vesta": 11426906929455361843568202299992114520848200991084027513389447476559454104162}


@pytest.fixture(scope="module")
def expected():
    return json.load(open(GOLDEN))


def compress(orc, base_fid, pt_mont) -> bytes:
    """affine Montgomery point -> 33-byte ark compressed form (utils/src/serialization.rs:65-84), with Python integers"""
    if not np.any(pt_mont):
        return bytes(32) + b"\x40"
    m = orc.MODULUS[base_fid]
    x, y = orc.limbs_to_ints(orc.from_mont(base_fid, np.ascontiguousarray(pt_mont).reshape(2, 4)))
    return x.to_bytes(32, "little") + bytes([0x80 if y > m - y else 0])


def msgpack_points(points33, struct_prefix=b"") -> bytes:
    assert len(points33) < 16
    return struct_prefix + bytes([0x90 | len(points33)]) + b"".join(b"\xc4\x21" + p for p in points33)


def padded(raw: bytes, n: int) -> list:
    """test_generic_serialization_regression_serde writes into a zeroed buffer of the expected length (serialization.rs:208-216)"""
    assert len(raw) <= n
    return list(raw) + [0] * (n - len(raw))


def test_trusted_setup_srs_bytes(orc, expected, pallas_srs, vesta_srs):
    rng = StdRng(bytes(32))
    for name, key, G, scalar_mod in (("vesta", "srs_vesta_trusted_setup_depth8", vesta_srs, orc.FP_MODULUS),
                                     ("pallas", "srs_pallas_trusted_setup_depth8", pallas_srs, orc.FQ_MODULUS)):
        x = orc.fe_int(G.scalar, np.array(rng.field_mont_limbs(scalar_mod), dtype=np.uint64))     # td = F::rand(rng)
        gen = orc.to_mont(G.base, orc.ints_to_limbs([1, GENERATOR_Y[name]])).reshape(8)
        pts, x_pow = [], 1
        for _ in range(8):
            pts.append(compress(orc, G.base, orc.scalar_mul(G.cid, gen, x_pow) if x_pow != 1 else gen))
            x_pow = x_pow * x % scalar_mod
        h = compress(orc, G.base, G.mont_points(G.h_xy_canon)[0])     # the same "srs_misc" blinder as SRS::create (ipa.rs:531-538)
        raw = b"\x92" + msgpack_points(pts) + b"\xc4\x21" + h
        assert padded(raw, len(expected[key])) == expected[key], name


def polycomm_inputs(orc):
    """poly = DensePolynomial::<Fp>::rand(300, rng) (301 coefficients), then one blinder per chunk: Montgomery limbs"""
    rng = StdRng(bytes(32))
    coeffs = np.array([rng.field_mont_limbs(orc.FP_MODULUS) for _ in range(301)], dtype=np.uint64)
    blinders = np.array([rng.field_mont_limbs(orc.FP_MODULUS) for _ in range(6)], dtype=np.uint64)
    return coeffs, blinders


def test_polycomm_bytes_with_the_oracle(orc, expected, vesta_srs):
    G = vesta_srs
    coeffs, blinders = polycomm_inputs(orc)
    g, h = G.g[:128], G.mont_points(G.h_xy_canon)[0]
    chunks = [orc.msm_mont(G.cid, g[: min(128, 301 - 128 * j)], coeffs[128 * j: 128 * (j + 1)]) for j in range(3)]
    chunks += [np.zeros(8, dtype=np.uint64)] * 3                       # padded with G::zero() up to num_chunks (ipa.rs:678-680)
    masked = []
    for c, b in zip(chunks, blinders):
        bh = orc.scalar_mul(G.cid, h, orc.fe_int(G.scalar, b))
        masked.append(compress(orc, G.base, orc.affine_add(G.cid, c, bh) if np.any(c) else bh))
    raw = msgpack_points(masked, struct_prefix=b"\x91")
    key = "polycomm_vesta_srs128_deg300_chunks6"
    assert padded(raw, len(expected[key])) == expected[key]


def test_poseidon_restatement_matches_the_reference_vectors():
    """poseidon/tests/test_vectors/kimchi.json (Fp, PlonkSpongeConstantsKimchi): pins tests/kimchi_transcript.py's permutation"""
    from kimchi_transcript import PARAMS, ArithmeticSponge
    for v in PARAMS["fp_hash_vectors"]:
        sp = ArithmeticSponge("fp")
        sp.absorb([int.from_bytes(bytes.fromhex(x), "little") for x in v["input"]])
        assert sp.squeeze().to_bytes(32, "little").hex() == v["output"]


def test_opening_proof_bytes_with_the_oracle(orc, expected, vesta_srs):
    """ser_regression_canonical_opening_proof (commitment.rs:388-443): the whole of SRS::open — combine_polys, the sponge, the
    group map, 7 folding rounds, delta, z1, z2, sg — replayed with the oracle doing the group arithmetic."""
    from open_replay import OracleRounds, first_opening_proof_bytes
    raw = first_opening_proof_bytes(orc, vesta_srs, lambda g, a, b: OracleRounds(orc, g, a, b))
    key = "opening_proof_vesta_srs128"
    assert padded(raw, len(expected[key])) == expected[key]
