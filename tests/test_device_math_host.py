"""The DEVICE arithmetic (proof_systems_b200/csrc/field.cuh, curve.cuh), compiled for the host with the PTX
carry-chain primitives emulated one instruction at a time, against the CPU oracle.  This checks the exact
mad.lo.cc / madc.hi.cc sequences and the XYZZ group law without a GPU; the same checks run on the device in
tests/test_gpu_field.py.
"""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "host_math.cpp")
SO = os.path.join(HERE, "host_math", "libhost_math.so")
CSRC = os.path.join(os.path.dirname(HERE), "proof_systems_b200", "csrc")


@pytest.fixture(scope="module")
def hm():
    deps = [SRC, os.path.join(CSRC, "field.cuh"), os.path.join(CSRC, "curve.cuh"), os.path.join(CSRC, "ntt_butterfly.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return ctypes.CDLL(SO)


def p32(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


def run_bin(hm, name, fid, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    r = np.empty_like(a)
    getattr(hm, name)(fid, p32(a.view(np.uint32)), p32(b.view(np.uint32)), p32(r.view(np.uint32)), ctypes.c_size_t(a.shape[0]))
    return r


def run_un(hm, name, fid, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    r = np.empty_like(a)
    getattr(hm, name)(fid, p32(a.view(np.uint32)), p32(r.view(np.uint32)), ctypes.c_size_t(a.shape[0]))
    return r


def sample_values(orc, fid, count, seed):
    m = orc.MODULUS[fid]
    rng = random.Random(seed)
    sp = [0, 1, 2, m - 1, m - 2, (1 << 254), (1 << 254) - 1, (1 << 32) - 1, 1 << 32, (1 << 224), m >> 1,
          (1 << 255) % m, 0xFFFFFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF, (m - 1) - 0xFFFFFFFF]
    vals = sp + [rng.randrange(m) for _ in range(count - len(sp))]
    return vals


@pytest.mark.parametrize("fid", [0, 1])
def test_constants(hm, orc, fid):
    one = np.empty(4, dtype=np.uint64); r2 = np.empty(4, dtype=np.uint64); root = np.empty(4, dtype=np.uint64)
    hm.hm_consts(fid, p32(one.view(np.uint32)), p32(r2.view(np.uint32)), p32(root.view(np.uint32)))
    assert np.array_equal(one, orc.const(fid, "one"))
    assert np.array_equal(r2, orc.const(fid, "r2"))
    assert np.array_equal(root, orc.two_adic_root(fid))


@pytest.mark.parametrize("fid", [0, 1])
def test_field_ops_vs_ints(hm, orc, fid):
    m = orc.MODULUS[fid]
    vals = sample_values(orc, fid, 300, 77 + fid)
    rng = random.Random(5)
    A = [rng.choice(vals) for _ in range(4000)] + vals
    B = [rng.choice(vals) for _ in range(4000)] + vals[::-1]
    a = orc.ints_to_limbs(A); b = orc.ints_to_limbs(B)
    Rinv = pow(1 << 256, -1, m)
    got = orc.limbs_to_ints(run_bin(hm, "hm_mul", fid, a, b))
    assert got == [x * y * Rinv % m for x, y in zip(A, B)]
    assert orc.limbs_to_ints(run_bin(hm, "hm_add", fid, a, b)) == [(x + y) % m for x, y in zip(A, B)]
    assert orc.limbs_to_ints(run_bin(hm, "hm_sub", fid, a, b)) == [(x - y) % m for x, y in zip(A, B)]
    assert orc.limbs_to_ints(run_un(hm, "hm_neg", fid, a)) == [(-x) % m for x in A]
    assert orc.limbs_to_ints(run_un(hm, "hm_to_mont", fid, a)) == [x * (1 << 256) % m for x in A]
    assert orc.limbs_to_ints(run_un(hm, "hm_from_mont", fid, a)) == [x * Rinv % m for x in A]
    small = orc.ints_to_limbs(vals[:40])
    inv = orc.limbs_to_ints(run_un(hm, "hm_inv", fid, orc.to_mont(fid, small)))
    assert inv == [(pow(v, -1, m) * (1 << 256)) % m if v else 0 for v in vals[:40]]


def xyzz_of(orc, cid, aff, z_int):
    """re-randomise an affine point into XYZZ with Z = z"""
    f = orc.BASE_FIELD[cid]
    if not np.any(aff):
        return np.zeros(16, dtype=np.uint64)
    z = orc.fe(f, z_int)
    zz = orc.fe_mul(f, z, z); zzz = orc.fe_mul(f, zz, z)
    return np.concatenate([orc.fe_mul(f, aff[:4], zz), orc.fe_mul(f, aff[4:], zzz), zz, zzz])


@pytest.mark.parametrize("name", ["pallas_srs", "vesta_srs"])
def test_xyzz_group_law_vs_oracle(hm, orc, request, name):
    srs = request.getfixturevalue(name)
    cid = srs.cid
    g = srs.g
    f = srs.base
    zero_aff = np.zeros(8, dtype=np.uint64)

    def neg(p):
        q = p.copy(); q[4:] = orc.fe_sub(f, np.zeros(4, dtype=np.uint64), p[4:]); return q

    def op(o, p_xyzz, q):
        oa = np.empty(8, dtype=np.uint64); ox = np.empty(16, dtype=np.uint64)
        hm.hm_curve_op(cid, o, p32(np.ascontiguousarray(p_xyzz).view(np.uint32)), p32(np.ascontiguousarray(q).view(np.uint32)),
                       p32(oa.view(np.uint32)), p32(ox.view(np.uint32)))
        return oa

    cases = [(g[0], g[1]), (g[2], g[2]), (g[3], neg(g[3])), (zero_aff, g[4]), (g[5], zero_aff), (zero_aff, zero_aff),
             (g[6], g[7]), (g[100], g[2000])]
    for k, (p, q) in enumerate(cases):
        want = orc.affine_add(cid, p, q)
        for z in (1, 7, 123456789123456789):
            px = xyzz_of(orc, cid, p, z)
            assert np.array_equal(op(0, px, q), want), ("madd", k, z)
            for z2 in (1, 99):
                qx = xyzz_of(orc, cid, q, z2)
                assert np.array_equal(op(1, px, qx), want), ("add", k, z, z2)
        want2 = orc.affine_add(cid, p, p)
        assert np.array_equal(op(2, xyzz_of(orc, cid, p, 5), zero_aff), want2), ("dbl", k)
    # a long accumulation: sum of 300 generators == MSM with all-one scalars
    n = 300
    out = np.empty(8, dtype=np.uint64)
    pts = np.ascontiguousarray(g[:n])
    hm.hm_sum_affine(cid, p32(pts.view(np.uint32)), ctypes.c_size_t(n), p32(out.view(np.uint32)))
    ones = np.zeros((n, 4), dtype=np.uint64); ones[:, 0] = 1
    assert np.array_equal(out, orc.msm(cid, pts, ones))


@pytest.mark.parametrize("fid", [0, 1])
@pytest.mark.parametrize("radix", [2, 4])
def test_ntt_tile_layer_schedule_vs_oracle(hm, orc, fid, radix):
    """The butterflies and index maps k_ntt_pass executes (csrc/ntt_butterfly.cuh: radix-2 layers and radix-2^2 units) on one
    column, every log_s the tile pass uses: the bit-reversed result must be the oracle's forward NTT of that size."""
    f = orc.FP if fid == 0 else orc.FQ
    m = orc.MODULUS[f]
    w1024 = orc.fe_int(f, orc.root_of_unity(f, 10))
    small = orc.to_mont(f, orc.ints_to_limbs([pow(w1024, i, m) for i in range(512)]))
    for log_s in range(0, 11):
        S = 1 << log_s
        a = orc.to_mont(f, orc.random_scalars(f, S, seed=100 + log_s))
        buf = np.ascontiguousarray(a).copy()
        hm.hm_ntt_column(fid, p32(buf.view(np.uint32)), log_s, p32(small.view(np.uint32)), radix)
        rev = [int(format(k, f"0{log_s}b")[::-1], 2) if log_s else 0 for k in range(S)]
        got = buf[rev]                      # out[k] sits in row bitrev(k)
        assert np.array_equal(got, orc.ntt(f, a)), (log_s, radix)


@pytest.mark.parametrize("fid", [0, 1])
def test_ntt_dit_schedule_with_register_stage_vs_oracle(hm, orc, fid):
    """The schedule k_ntt_pass runs since round 2 (csrc/ntt.cu): bit-reversed load, the six layers of span <= 32 in REGISTERS —
    emulated here lane by lane with the very per-lane functions the kernel calls (ntt_lane_pre / ntt_lane_post / ntt_lane_last,
    csrc/ntt_butterfly.cuh), the warp shuffles replaced by array reads — then the shared-memory layers; natural order out.
    Every log_s the tile pass uses must give the oracle's forward NTT."""
    f = orc.FP if fid == 0 else orc.FQ
    m = orc.MODULUS[f]
    w1024 = orc.fe_int(f, orc.root_of_unity(f, 10))
    small = orc.to_mont(f, orc.ints_to_limbs([pow(w1024, i, m) for i in range(512)]))
    for log_s in range(0, 11):
        S = 1 << log_s
        a = orc.to_mont(f, orc.random_scalars(f, S, seed=200 + log_s))
        buf = np.ascontiguousarray(a).copy()
        hm.hm_ntt_column_dit(fid, p32(buf.view(np.uint32)), log_s, p32(small.view(np.uint32)))
        assert np.array_equal(buf, orc.ntt(f, a)), log_s


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_split_product_variants_are_bit_exact(orc, k):
    """field.cuh's ZK_MUL_PLAIN_PER_ROW = k (k products per row as stand-alone wide multiplies + carry adds: the pipe-balancing
    experiment of DESIGN.md) computes the same Montgomery product as the shipped form, on edge values and random ones."""
    so = os.path.join(HERE, "host_math", f"libhost_math_k{k}.so")
    deps = [SRC, os.path.join(CSRC, "field.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-DZK_MUL_PLAIN_PER_ROW={k}", "-o", so, SRC])
    lib = ctypes.CDLL(so)
    for fid in (0, 1):
        f = orc.FP if fid == 0 else orc.FQ
        m = orc.MODULUS[f]
        vals = sample_values(orc, f, 400, seed=70 + k)
        a = orc.ints_to_limbs(vals)
        b = orc.ints_to_limbs(vals[::-1])
        got = run_bin(lib, "hm_mul", fid, a, b)
        rinv = pow(1 << 256, -1, m)
        assert orc.limbs_to_ints(got) == [x * y * rinv % m for x, y in zip(vals, vals[::-1])]
