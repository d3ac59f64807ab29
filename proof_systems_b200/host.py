"""Host-side mirrors of the reference interfaces on the MSM/NTT path, over the C ABI (srs.cu holds the policy code).

Names, argument meaning and error behaviour follow the reference so that the parity tests read like its own:
  poly_commitment::SRS / ipa::SRS<G>                  poly-commitment/src/lib.rs:61-241, ipa.rs:56-75,596-800
  poly_commitment::PolyComm<C>{chunks}                poly-commitment/src/commitment.rs:47-50
  poly_commitment::error::CommitmentError             poly-commitment/src/error.rs:3-9
  ark_poly::Radix2EvaluationDomain<F>                 used as `D` in kimchi/src/prover.rs:39-42, circuits/domains.rs:24-33
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from ._lib import BASE_FIELD, SCALAR_FIELD, Context, ZkError, _np_u64, _ptr, _u64p, check, lib


class BlindersDontMatch(ValueError):
    """CommitmentError::BlindersDontMatch(blinders_len, commitment_len)"""


@dataclass
class PolyComm:
    """chunks: uint64 [k, 8] affine points (identity = zeros)"""
    chunks: np.ndarray

    def __len__(self):
        return self.chunks.shape[0]


class SRS:
    """ipa::SRS<G>{g, h, lagrange_bases} with g (and every registered Lagrange basis) resident on the device."""

    def __init__(self, ctx: Context, curve: int, g, h, window_bits: int = -1):
        self.ctx, self.curve = ctx, curve
        self.g = _np_u64(g, (8,))
        self.h = np.ascontiguousarray(h, dtype=np.uint64).reshape(8)
        self._h = ctypes.c_void_p()
        check(lib().zk_srs_create(ctx._h, curve, _ptr(self.g), self.g.shape[0], self.h.ctypes.data_as(_u64p), window_bits, ctypes.byref(self._h)))

    @classmethod
    def from_file(cls, ctx: Context, curve: int, path: str, window_bits: int = -1, lagrange: bool = True) -> "SRS":
        """Load srs/{pallas,vesta}.srs or srs/test_{pallas,vesta}.srs (precomputed_srs.rs:76-91 get_srs / get_srs_test): the points
        are decoded on the device; the Lagrange bases stored in a test file populate the cache (single-chunk bases only)."""
        from . import srs_file
        f = srs_file.read_srs(path)
        dec = ctx.decompress_points if f.compressed else ctx.points_from_uncompressed
        srs = cls(ctx, curve, dec(curve, f.g), dec(curve, f.h.reshape(1, -1))[0], window_bits)
        if lagrange:
            for n, basis in f.lagrange_bases.items():
                if basis.shape[1] == 1:
                    srs.add_lagrange_basis(n, ctx.points_from_uncompressed(curve, basis[:, 0]))
        return srs

    def close(self):
        if getattr(self, "_h", None):
            lib().zk_srs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass

    # fn max_poly_size(&self) -> usize  /  fn size(&self)
    def max_poly_size(self) -> int:
        return int(lib().zk_srs_max_poly_size(self._h))

    size = max_poly_size

    def blinding_commitment(self) -> np.ndarray:
        return self.h

    def add_lagrange_basis(self, domain_size: int, basis, window_bits: int = -1):
        """Populate the cache behind get_lagrange_basis(domain) (ipa.rs:780-795) with a precomputed basis."""
        b = _np_u64(basis, (8,))
        assert b.shape[0] == domain_size * self.lagrange_basis_chunks(domain_size)      # chunk-major for domains larger than the SRS
        check(lib().zk_srs_add_lagrange_basis(self._h, domain_size, _ptr(b), window_bits))

    def lagrange_basis_chunks(self, domain_size: int) -> int:
        """chunks per basis element: ceil(domain / |g|) (ipa.rs:1145)"""
        return int(lib().zk_srs_lagrange_basis_chunks(self._h, domain_size))

    # fn get_lagrange_basis_from_domain_size(&self, domain_size: usize) -> &Vec<PolyComm<G>>
    def get_lagrange_basis_from_domain_size(self, domain_size: int, window_bits: int = -1) -> np.ndarray:
        """Computes (once) the basis on the device — SRS::lagrange_basis, ipa.rs:1065-1172 — and returns it: [n, 8] affine, or
        [chunks, n, 8] (chunk-major: PolyComm i of the reference is out[:, i]) when the domain is larger than the SRS."""
        check(lib().zk_srs_lagrange_basis(self._h, domain_size, window_bits))
        chunks = self.lagrange_basis_chunks(domain_size)
        out = np.empty((chunks * domain_size, 8), dtype=np.uint64)
        check(lib().zk_srs_get_lagrange_basis(self._h, domain_size, out.ctypes.data_as(_u64p), chunks * domain_size))
        return out if chunks == 1 else out.reshape(chunks, domain_size, 8)

    # fn commit_non_hiding(&self, plnm: &DensePolynomial<F>, num_chunks: usize) -> PolyComm<G>
    def commit_non_hiding(self, coeffs, num_chunks: int) -> PolyComm:
        c = _np_u64(coeffs, (4,)) if len(coeffs) else np.zeros((0, 4), dtype=np.uint64)
        cap = max(num_chunks, (c.shape[0] + self.g.shape[0] - 1) // self.g.shape[0], 1)
        out = np.zeros((cap, 8), dtype=np.uint64)
        k = ctypes.c_size_t()
        check(lib().zk_srs_commit_non_hiding(self._h, _ptr(c) if c.size else None, c.shape[0], num_chunks,
                                             out.ctypes.data_as(_u64p), cap, ctypes.byref(k)))
        return PolyComm(out[: k.value])

    # fn commit_evaluations_non_hiding(&self, domain: D<F>, plnm: &Evaluations<F, D<F>>) -> PolyComm<G>
    def commit_evaluations_non_hiding(self, domain_size: int, evals) -> PolyComm:
        e = _np_u64(evals, (4,))
        out = np.zeros((max(1, self.lagrange_basis_chunks(domain_size)), 8), dtype=np.uint64)
        check(lib().zk_srs_commit_evaluations_non_hiding(self._h, domain_size, _ptr(e), e.shape[0], out.ctypes.data_as(_u64p)))
        return PolyComm(out)

    def commit_evaluations_non_hiding_batch(self, domain_size: int, evals) -> list:
        """[commit_evaluations_non_hiding(domain, e) for e in evals] in one call: evals uint64 [k, domain_size, 4].  The
        reference computes the witness commitments concurrently (kimchi/src/prover.rs:329-351); here the k MSMs share lanes."""
        e = np.ascontiguousarray(evals, dtype=np.uint64)
        if e.ndim != 3 or e.shape[1:] != (domain_size, 4):
            raise ValueError("expected uint64 [k, domain_size, 4]")
        out = np.zeros((e.shape[0], 8), dtype=np.uint64)
        check(lib().zk_srs_commit_evaluations_batch(self._h, domain_size, _ptr(e), e.shape[0], out.ctypes.data_as(_u64p)))
        return [PolyComm(out[j:j + 1]) for j in range(e.shape[0])]

    # fn mask_custom(&self, com: PolyComm<G>, blinders: &PolyComm<F>) -> Result<BlindedCommitment<G>, CommitmentError>
    def mask_custom(self, com: PolyComm, blinders) -> PolyComm:
        b = _np_u64(blinders, (4,))
        out = np.zeros_like(com.chunks)
        try:
            check(lib().zk_srs_mask_custom(self._h, _ptr(com.chunks), len(com), _ptr(b), b.shape[0], out.ctypes.data_as(_u64p)))
        except ZkError as e:
            if e.code == -4:
                raise BlindersDontMatch(b.shape[0], len(com)) from e
            raise
        return PolyComm(out)

    def commit_custom(self, coeffs, num_chunks: int, blinders) -> PolyComm:
        return self.mask_custom(self.commit_non_hiding(coeffs, num_chunks), blinders)

    def commit_evaluations_custom(self, domain_size: int, evals, blinders) -> PolyComm:
        return self.mask_custom(self.commit_evaluations_non_hiding(domain_size, evals), blinders)


class IndexCache:
    """A cached prover index (kimchi/src/cached_prover_index.rs:26-56, "MINAPK01") resident on the device: zk_index_cache_load parses the
    header and the section table and copies the payload as it lies in the file — raw Montgomery limbs are the device format."""

    def __init__(self, ctx: Context, image: bytes, expect_identifier: str | None = None):
        from ._lib import IndexHeader
        self.ctx = ctx
        self._image = np.frombuffer(image, dtype=np.uint8)          # keeps the bytes alive during the copy
        self._h = ctypes.c_void_p()
        ident = expect_identifier.encode() if expect_identifier is not None else None
        check(lib().zk_index_cache_load(ctx._h, ctypes.c_void_p(self._image.ctypes.data), self._image.size, ident, ctypes.byref(self._h)))
        hdr = IndexHeader()
        check(lib().zk_index_cache_header(self._h, ctypes.byref(hdr)))
        self.header = hdr

    @classmethod
    def from_file(cls, ctx: Context, path: str, expect_identifier: str | None = None) -> "IndexCache":
        import mmap
        with open(path, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        return cls(ctx, mm, expect_identifier)

    def section(self, tag: int):
        """(device pointer, element count, declared domain size) of a field-element section"""
        p, n, d = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_uint32()
        check(lib().zk_index_cache_section(self._h, tag, ctypes.byref(p), ctypes.byref(n), ctypes.byref(d)))
        return p.value or 0, n.value, d.value

    def close(self):
        if getattr(self, "_h", None):
            lib().zk_index_cache_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass


class OpeningProof:
    """poly_commitment::ipa::OpeningProof (ipa.rs:1175-1191): lr [rounds, 2, 8], delta [8], z1 [4], z2 [4], sg [8] — points affine,
    everything in Montgomery limbs"""

    def __init__(self, lr, delta, z1, z2, sg):
        self.lr, self.delta, self.z1, self.z2, self.sg = lr, delta, z1, z2, sg


def srs_open(srs, plnms, elm, polyscale, evalscale, rng_scalars, u_base, round_challenge, final_challenge) -> OpeningProof:
    """SRS::open (ipa.rs:823-1061) through zk_srs_open.
    plnms: list of (data [len, 4] Montgomery numpy array OR (device_ptr, len), domain_size (0 = coefficients), blinders [k, 4]);
    elm [m, 4]; rng_scalars [2 * rounds + 2, 4] = rand_l, rand_r per round, then d, r_delta; the three callables are the caller's
    transcript: u_base(cip [4]) -> U [8]; round_challenge(i, l [8], r [8]) -> u [4]; final_challenge(delta [8]) -> c [4]."""
    from ._lib import FINAL_CB, ROUND_CB, U_BASE_CB, OpenPoly, OpenTranscript
    keep, arr = [], (OpenPoly * max(1, len(plnms)))()
    for i, (data, dom, blinders) in enumerate(plnms):
        if isinstance(data, tuple):
            ptr, ln = data
        else:
            d = _np_u64(data, (4,))
            keep.append(d)
            ptr, ln = d.ctypes.data, d.shape[0]
        bl = _np_u64(blinders, (4,))
        keep.append(bl)
        arr[i] = OpenPoly(ptr, ln, dom, bl.ctypes.data, bl.shape[0])
    elm = _np_u64(elm, (4,))
    rng = _np_u64(rng_scalars, (4,))
    ps = np.ascontiguousarray(polyscale, dtype=np.uint64).reshape(4)
    es = np.ascontiguousarray(evalscale, dtype=np.uint64).reshape(4)
    errors = []

    def view(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def put(p, v, n):
        np.ctypeslib.as_array(p, shape=(n,))[:] = np.ascontiguousarray(v, dtype=np.uint64).reshape(n)

    def guard(fn):
        def wrapped(*a):
            try:
                fn(*a)
                return 0
            except Exception as e:           # never unwind through the C frames
                errors.append(e)
                return 1
        return wrapped

    cb_u = U_BASE_CB(guard(lambda user, cip, out: put(out, u_base(view(cip, 4)), 8)))
    cb_r = ROUND_CB(guard(lambda user, i, l, r, out: put(out, round_challenge(int(i), view(l, 8), view(r, 8)), 4)))
    cb_f = FINAL_CB(guard(lambda user, d, out: put(out, final_challenge(view(d, 8)), 4)))
    tr = OpenTranscript(None, cb_u, cb_r, cb_f)
    rounds_cap = 64
    lr = np.zeros((rounds_cap, 2, 8), dtype=np.uint64)
    delta, sg = np.zeros(8, dtype=np.uint64), np.zeros(8, dtype=np.uint64)
    z1, z2 = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    rounds = ctypes.c_size_t(0)
    rc = lib().zk_srs_open(srs._h, arr, len(plnms), _ptr(elm), elm.shape[0], _ptr(ps), _ptr(es), _ptr(rng), rng.shape[0], ctypes.byref(tr),
                           _ptr(lr), rounds_cap, ctypes.byref(rounds), _ptr(delta), _ptr(z1), _ptr(z2), _ptr(sg))
    if errors:
        raise errors[0]
    check(rc)
    return OpeningProof(lr[: rounds.value].copy(), delta, z1, z2, sg)


class IpaRounds:
    """The folding loop of SRS::open (poly-commitment/src/ipa.rs:929-1007) with a and b resident on the device and the bases
    taken from the resident SRS table (`bases` = ctx.upload_bases(curve, srs.g)); see csrc/ipa.cu.

    The caller owns the sponge: per round it takes `lr()` -> (<a_hi,g_lo>, <a_lo,g_hi>, <a_hi,b_lo>, <a_lo,b_hi>), finishes
    L and R with its blinders (rand_l * h + <a_hi,b_lo> * u_base, ipa.rs:943-961), absorbs them, squeezes u and calls
    `fold(u, u_inv)`.  After log2(n) rounds `state()` is (a0, b0) and `sg()` the folded base g0."""

    def __init__(self, ctx: Context, bases, a_mont, b_mont):
        a, b = _np_u64(a_mont, (4,)), _np_u64(b_mont, (4,))
        n = 1 << max(1, (len(bases) - 1).bit_length())          # ipa.rs:848-850: padded_length
        if a.shape[0] > n or b.shape[0] > n:
            raise ValueError("a and b must not be longer than the padded SRS")
        pad = lambda v: np.concatenate([v, np.zeros((n - v.shape[0], 4), dtype=np.uint64)]) if v.shape[0] < n else v
        a, b = pad(a), pad(b)                                    # ipa.rs:858-862: a padded with zeros
        self.ctx, self.curve, self.bases = ctx, bases.curve, bases
        self._h = ctypes.c_void_p()
        check(lib().zk_ipa_begin(ctx._h, bases._h, _ptr(a), _ptr(b), n, ctypes.byref(self._h)))

    def __len__(self):
        return int(lib().zk_ipa_len(self._h))

    def lr(self):
        l, r = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
        ipl, ipr = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        check(lib().zk_ipa_round_lr(self._h, l.ctypes.data_as(_u64p), r.ctypes.data_as(_u64p), ipl.ctypes.data_as(_u64p), ipr.ctypes.data_as(_u64p)))
        return l, r, ipl, ipr

    def fold(self, u_mont, u_inv_mont):
        u = np.ascontiguousarray(u_mont, dtype=np.uint64).reshape(4)
        ui = np.ascontiguousarray(u_inv_mont, dtype=np.uint64).reshape(4)
        check(lib().zk_ipa_round_fold(self._h, u.ctypes.data_as(_u64p), ui.ctypes.data_as(_u64p)))

    def state(self):
        n = len(self)
        a, b = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 4), dtype=np.uint64)
        check(lib().zk_ipa_read(self._h, _ptr(a), _ptr(b), n, None))
        return a, b

    def sg(self) -> np.ndarray:
        """g0 after the last fold, Jacobian [12]"""
        out = np.zeros(12, dtype=np.uint64)
        check(lib().zk_ipa_read(self._h, None, None, 0, out.ctypes.data_as(_u64p)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().zk_ipa_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass


class ExprProgram:
    """Builder of the RPN programs zk_expr_eval_dev runs — kimchi's `PolishToken` list (kimchi/src/circuits/expr.rs:819-836) with the
    Challenge / Constant terms already literal.  Methods are the reference's token names; `evaluations` is Expr::evaluations
    (expr.rs:1938-1990) on the device.  Opcode numbers are include/zkb200.h's ZK_EXPR_*."""
    CONST, CELL, DUP, POW, ADD, MUL, SUB, STORE, LOAD = range(9)

    def __init__(self):
        self.tokens: list[tuple[int, int]] = []
        self.constants: list[np.ndarray] = []
        self.n_cached = 0

    def literal(self, x_mont):
        self.constants.append(np.ascontiguousarray(x_mont, dtype=np.uint64).reshape(4))
        self.tokens.append((self.CONST, len(self.constants) - 1)); return self

    def cell(self, col: int, next_row: bool = False):
        self.tokens.append((self.CELL, col | (0x80000000 if next_row else 0))); return self

    def dup(self): self.tokens.append((self.DUP, 0)); return self
    def pow(self, n: int): self.tokens.append((self.POW, n)); return self
    def add(self): self.tokens.append((self.ADD, 0)); return self
    def mul(self): self.tokens.append((self.MUL, 0)); return self
    def sub(self): self.tokens.append((self.SUB, 0)); return self

    def store(self) -> int:
        """Store: the top of the stack also goes to the next cache slot; returns the slot for load()"""
        self.tokens.append((self.STORE, 0)); self.n_cached += 1; return self.n_cached - 1

    def load(self, slot: int): self.tokens.append((self.LOAD, slot)); return self

    def evaluations(self, ctx: Context, field: int, cols, out_len: int, out_domain_mult: int, d_out: int, accumulate: bool = False):
        """cols: [(device pointer, len, domain_mult)] in the order the program's cell() indices refer to"""
        ctx.expr_eval_dev(field, self.tokens, np.array(self.constants, dtype=np.uint64).reshape(-1, 4), cols, out_len, out_domain_mult, d_out, accumulate)


class Radix2EvaluationDomain:
    """ark_poly::Radix2EvaluationDomain::<F>::new(size) on the device.  `field` is ZK_FP or ZK_FQ."""

    def __init__(self, ctx: Context, field: int, size: int):
        if size <= 0:
            raise ValueError("domain size must be positive")
        log_n = (size - 1).bit_length()  # new(n) rounds up to the next power of two
        self.ctx, self.field, self.log_size_of_group, self.size = ctx, field, log_n, 1 << log_n

    def _run(self, a, inverse, coset):
        a = _np_u64(a, (4,))
        if a.shape[0] > self.size:
            raise ValueError("more coefficients than the domain size")  # ark reduces mod X^n - 1; callers on the path never do
        buf = np.zeros((self.size, 4), dtype=np.uint64)
        buf[: a.shape[0]] = a
        return self.ctx.ntt(self.field, buf, inverse=inverse, coset=coset, in_len=a.shape[0] if not inverse else 0)

    def fft(self, coeffs) -> np.ndarray:
        return self._run(coeffs, False, False)

    def ifft(self, evals) -> np.ndarray:
        return self._run(evals, True, False)

    def coset_fft(self, coeffs) -> np.ndarray:
        return self._run(coeffs, False, True)

    def coset_ifft(self, evals) -> np.ndarray:
        return self._run(evals, True, True)
