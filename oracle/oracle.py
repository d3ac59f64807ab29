"""ctypes front-end of the CPU oracle (oracle/pasta_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package (proof_systems_b200) never does.

Conventions (same as the C-ABI of the product, include/zkb200.h):
  field elements  numpy uint64 [..., 4]   little-endian limbs, Montgomery form unless stated
  affine points   numpy uint64 [..., 8]   x || y, identity = all zero
  field_id        0 = Fp, 1 = Fq ;  curve_id 0 = Pallas (base Fp, scalars Fq), 1 = Vesta (base Fq, scalars Fp)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpasta_oracle.so")

FP, FQ = 0, 1
PALLAS, VESTA = 0, 1

FP_MODULUS = 28948022309329048855892746252171976963363056481941560715954676764349967630337
FQ_MODULUS = 28948022309329048855892746252171976963363056481941647379679742748393362948097
MODULUS = {FP: FP_MODULUS, FQ: FQ_MODULUS}
BASE_FIELD = {PALLAS: FP, VESTA: FQ}
SCALAR_FIELD = {PALLAS: FQ, VESTA: FP}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("pasta_oracle.c", "field_impl.h", "curve_impl.h", "ntt_impl.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        sz = ctypes.c_size_t
        i = ctypes.c_int
        u = ctypes.c_uint
        L.orc_max_threads.restype = i
        L.orc_fe_mul.argtypes = [i, u64p, u64p, u64p]
        L.orc_fe_add.argtypes = [i, u64p, u64p, u64p]
        L.orc_fe_sub.argtypes = [i, u64p, u64p, u64p]
        L.orc_fe_inv.argtypes = [i, u64p, u64p]
        L.orc_fe_sqrt.argtypes = [i, u64p, u64p]
        L.orc_fe_sqrt.restype = i
        L.orc_fe_pow.argtypes = [i, u64p, u64p, u64p]
        L.orc_fe_to_mont.argtypes = [i, u64p, u64p, sz]
        L.orc_fe_from_mont.argtypes = [i, u64p, u64p, sz]
        L.orc_fe_root_of_unity.argtypes = [i, u, u64p]
        L.orc_fe_two_adic_root.argtypes = [i, u64p]
        L.orc_fe_modulus.argtypes = [i, u64p]
        L.orc_fe_r2.argtypes = [i, u64p]
        L.orc_fe_one.argtypes = [i, u64p]
        L.orc_ntt.argtypes = [i, u64p, u, i, i, i]
        L.orc_dft_naive.argtypes = [i, u64p, u64p, u, i]
        L.orc_on_curve.argtypes = [i, u64p]
        L.orc_on_curve.restype = i
        L.orc_affine_add.argtypes = [i, u64p, u64p, u64p]
        L.orc_jac_add.argtypes = [i, u64p, u64p, u64p]
        L.orc_jac_to_affine.argtypes = [i, u64p, u64p]
        L.orc_scalar_mul.argtypes = [i, u64p, u64p, u64p]
        L.orc_decompress.argtypes = [i, u8p, u64p, sz, i]
        L.orc_decompress.restype = sz
        L.orc_msm.argtypes = [i, u64p, u64p, sz, i, i, u64p, u64p]
        L.orc_msm_mont.argtypes = [i, u64p, u64p, sz, i, u64p]
        L.orc_msm_split2.argtypes = [i, u64p, u64p, sz, i, u64p]
        L.orc_group_intt.argtypes = [i, u64p, u64p, u, i]
        L.orc_extend_bases.argtypes = [i, u64p, sz, u64p, sz]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


# ---------------------------------------------------------------- int <-> limb helpers
def int_to_limbs(x: int) -> np.ndarray:
    return np.array([(x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], dtype=np.uint64)


def limbs_to_int(a) -> int:
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[k]) << (64 * k) for k in range(len(a)))


def ints_to_limbs(xs) -> np.ndarray:
    out = np.empty((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i] = int_to_limbs(x)
    return out


def limbs_to_ints(a: np.ndarray) -> list:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [limbs_to_int(r) for r in a]


def bytes_le_to_limbs(b: bytes) -> np.ndarray:
    """n*32 bytes of LE canonical integers -> [n,4] uint64."""
    return np.frombuffer(b, dtype="<u8").reshape(-1, 4).copy()


# ---------------------------------------------------------------- field
def to_mont(fid: int, a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    r = np.empty_like(a)
    lib().orc_fe_to_mont(fid, _p(a), _p(r), a.size // 4)
    return r


def from_mont(fid: int, a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    r = np.empty_like(a)
    lib().orc_fe_from_mont(fid, _p(a), _p(r), a.size // 4)
    return r


def fe(fid: int, x: int) -> np.ndarray:
    """canonical int -> Montgomery limbs"""
    return to_mont(fid, int_to_limbs(x % MODULUS[fid]))


def fe_int(fid: int, a) -> int:
    """Montgomery limbs -> canonical int"""
    return limbs_to_int(from_mont(fid, np.asarray(a, dtype=np.uint64)))


def _binop(name):
    def f(fid, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        r = np.empty(4, dtype=np.uint64)
        getattr(lib(), name)(fid, _p(a), _p(b), _p(r))
        return r
    return f


fe_mul = _binop("orc_fe_mul")
fe_add = _binop("orc_fe_add")
fe_sub = _binop("orc_fe_sub")


def fe_inv(fid, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    r = np.empty(4, dtype=np.uint64)
    lib().orc_fe_inv(fid, _p(a), _p(r))
    return r


def fe_sqrt(fid, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    r = np.empty(4, dtype=np.uint64)
    ok = lib().orc_fe_sqrt(fid, _p(a), _p(r))
    return r if ok else None


def fe_pow(fid, a, e: int):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    r = np.empty(4, dtype=np.uint64)
    ee = int_to_limbs(e)
    lib().orc_fe_pow(fid, _p(a), _p(ee), _p(r))
    return r


def root_of_unity(fid, log_n) -> np.ndarray:
    r = np.empty(4, dtype=np.uint64)
    lib().orc_fe_root_of_unity(fid, log_n, _p(r))
    return r


def two_adic_root(fid) -> np.ndarray:
    r = np.empty(4, dtype=np.uint64)
    lib().orc_fe_two_adic_root(fid, _p(r))
    return r


def const(fid, which) -> np.ndarray:
    r = np.empty(4, dtype=np.uint64)
    getattr(lib(), {"modulus": "orc_fe_modulus", "r2": "orc_fe_r2", "one": "orc_fe_one"}[which])(fid, _p(r))
    return r


# ---------------------------------------------------------------- NTT
def ntt(fid: int, data: np.ndarray, inverse: bool = False, coset: bool = False, threads: int = 0) -> np.ndarray:
    """Radix2EvaluationDomain::{fft,ifft}_in_place on a copy; data [n,4] Montgomery, n a power of two."""
    a = np.array(data, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    lib().orc_ntt(fid, _p(a), log_n, int(inverse), int(coset), threads)
    return a


def perm_quot(fid: int, w, z, sigma, zkpm, beta, gamma, alpha0, shifts, next_shift: int = 8, threads: int = 0) -> np.ndarray:
    """permutation part of kimchi's quotient over d8 (permutation.rs:223-357): w, sigma [7, m, 4]; z, zkpm [m, 4]; beta, gamma, alpha0 [4];
    shifts [7, 4]; all Montgomery.  Returns [m, 4]."""
    w = np.ascontiguousarray(w, dtype=np.uint64).reshape(7, -1, 4)
    sigma = np.ascontiguousarray(sigma, dtype=np.uint64).reshape(7, -1, 4)
    m = w.shape[1]
    log_m = m.bit_length() - 1
    assert 1 << log_m == m and sigma.shape[1] == m
    z = np.ascontiguousarray(z, dtype=np.uint64).reshape(m, 4)
    zkpm = np.ascontiguousarray(zkpm, dtype=np.uint64).reshape(m, 4)
    out = np.empty((m, 4), dtype=np.uint64)
    c = lambda a, k: np.ascontiguousarray(a, dtype=np.uint64).reshape(k)
    lib().orc_perm_quot(fid, _p(w), ctypes.c_size_t(m), _p(z), _p(sigma), ctypes.c_size_t(m), _p(zkpm), _p(c(beta, 4)), _p(c(gamma, 4)), _p(c(alpha0, 4)),
                        _p(c(shifts, 28)), next_shift, log_m, _p(out), threads)
    return out


def expr_eval(fid: int, ops, args, literals, cols, out_len: int, acc: np.ndarray | None = None, threads: int = 0) -> np.ndarray:
    """PolishToken::evaluate (expr.rs:856-940) at every index of a domain of out_len points.  ops/args: the program (opcodes
    0 literal, 1 cell, 2 dup, 3 pow, 4 add, 5 mul, 6 sub, 7 store, 8 load); literals [k, 4] Montgomery; cols: list of
    (evals [len, 4] Montgomery, domain_mult).  acc: accumulate into a copy of this array.  Raises ValueError on the reference's
    failure modes (empty stack, final stack != 1, index out of range)."""
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    args = np.ascontiguousarray(args, dtype=np.uint32)
    lit = np.ascontiguousarray(literals, dtype=np.uint64).reshape(-1, 4)
    arrs = [np.ascontiguousarray(c[0], dtype=np.uint64).reshape(-1, 4) for c in cols]
    ptrs = (ctypes.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
    lens = np.array([a.shape[0] for a in arrs] or [0], dtype=np.uint64)
    mult = np.array([c[1] for c in cols] or [0], dtype=np.uint32)
    out = np.zeros((out_len, 4), dtype=np.uint64) if acc is None else np.ascontiguousarray(acc, dtype=np.uint64).reshape(out_len, 4).copy()
    rc = lib().orc_expr_eval(fid, ops.ctypes.data_as(ctypes.c_void_p), args.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ops)), _p(lit) if lit.size else None,
                             ctypes.c_size_t(lit.shape[0]), ptrs, lens.ctypes.data_as(ctypes.c_void_p), mult.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(arrs)),
                             ctypes.c_uint64(out_len), int(acc is not None), _p(out), threads)
    if rc != 0:
        raise ValueError("malformed RPN program")
    return out


def divide_by_vanishing(fid: int, coeffs, log_n: int):
    """DensePolynomial::divide_by_vanishing_poly(d1) -> (quotient [max(len - n, 0), 4], remainder [n, 4]); coefficients Montgomery"""
    f = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    n = 1 << log_n
    quot = np.zeros((max(f.shape[0] - n, 0), 4), dtype=np.uint64)
    rem = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_divide_by_vanishing(fid, _p(f), ctypes.c_size_t(f.shape[0]), log_n, _p(quot) if quot.size else None, _p(rem))
    return quot, rem


def dft_naive(fid: int, data: np.ndarray, inverse: bool = False) -> np.ndarray:
    a = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
    n = a.shape[0]
    log_n = n.bit_length() - 1
    out = np.empty_like(a)
    lib().orc_dft_naive(fid, _p(a), _p(out), log_n, int(inverse))
    return out


# ---------------------------------------------------------------- curve
def on_curve(cid, xy) -> bool:
    xy = np.ascontiguousarray(xy, dtype=np.uint64)
    return bool(lib().orc_on_curve(cid, _p(xy)))


def affine_add(cid, p, q) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.uint64)
    q = np.ascontiguousarray(q, dtype=np.uint64)
    r = np.empty(8, dtype=np.uint64)
    lib().orc_affine_add(cid, _p(p), _p(q), _p(r))
    return r


def jac_to_affine(cid, p) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.uint64)
    r = np.empty(8, dtype=np.uint64)
    lib().orc_jac_to_affine(cid, _p(p), _p(r))
    return r


def jac_add(cid, p, q) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.uint64)
    q = np.ascontiguousarray(q, dtype=np.uint64)
    r = np.empty(12, dtype=np.uint64)
    lib().orc_jac_add(cid, _p(p), _p(q), _p(r))
    return r


def scalar_mul(cid, p, k: int) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.uint64)
    kk = int_to_limbs(k)
    r = np.empty(8, dtype=np.uint64)
    lib().orc_scalar_mul(cid, _p(p), _p(kk), _p(r))
    return r


def decompress(cid, raw33: bytes, threads: int = 0) -> np.ndarray:
    n = len(raw33) // 33
    buf = np.frombuffer(raw33, dtype=np.uint8).copy()
    out = np.empty((n, 8), dtype=np.uint64)
    ok = lib().orc_decompress(cid, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _p(out), n, threads)
    if ok != n:
        raise ValueError(f"decompression failed for {n - ok} points")
    return out


def msm(cid, bases: np.ndarray, scalars: np.ndarray, algo: int = 0, threads: int = 0, want_jac: bool = False):
    """G::Group::msm_bigint(bases, scalars).into_affine(); scalars canonical [n,4]."""
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    aff = np.empty(8, dtype=np.uint64)
    jac = np.empty(12, dtype=np.uint64)
    lib().orc_msm(cid, _p(bases), _p(scalars), n, algo, threads, _p(jac), _p(aff))
    return (aff, jac) if want_jac else aff


def msm_split2(cid, bases, scalars, threads: int = 0) -> np.ndarray:
    """commit_non_hiding's `len == |g|` branch (ipa.rs:652-662): rayon::join of two half MSMs, then add."""
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    aff = np.empty(8, dtype=np.uint64)
    lib().orc_msm_split2(cid, _p(bases), _p(scalars), n, threads, _p(aff))
    return aff


def host_threads() -> int:
    """threads the CPU baseline may use: the cores this process is allowed on (torchrun pins OMP_NUM_THREADS=1)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def msm_mont(cid, bases, scalars_mont, threads: int = 0) -> np.ndarray:
    """G::Group::msm(bases, scalars).unwrap().into_affine(); scalars Montgomery."""
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    scalars_mont = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    assert bases.shape[0] == scalars_mont.shape[0]
    aff = np.empty(8, dtype=np.uint64)
    lib().orc_msm_mont(cid, _p(bases), _p(scalars_mont), bases.shape[0], threads, _p(aff))
    return aff


def extend_bases(cid, points: np.ndarray, n: int) -> np.ndarray:
    """Deterministic n-point base set from m fixture points (block k = block k-1 + rotated fixture; batch-affine)."""
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_extend_bases(cid, _p(pts), pts.shape[0], _p(out), n)
    return out


def group_intt(cid, points: np.ndarray, threads: int = 0) -> np.ndarray:
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
    n = pts.shape[0]
    log_n = n.bit_length() - 1
    out = np.empty_like(pts)
    lib().orc_group_intt(cid, _p(pts), _p(out), log_n, threads)
    return out


# ---------------------------------------------------------------- deterministic synthetic inputs
def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """count uint64 words of splitmix64 (SURVEY.md §8d input recipe), vectorised."""
    idx = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def random_scalars(fid: int, n: int, seed: int) -> np.ndarray:
    """n canonical field elements: 4 splitmix64 limbs, top limb masked to 62 bits (so the value is < 2^254 < m).
    Uniform over [0, 2^254), which covers all but a 2^-128 fraction of the field."""
    w = splitmix64_stream(seed, 4 * n).reshape(n, 4).copy()
    w[:, 3] &= np.uint64((1 << 62) - 1)
    return w
