"""RPN programs of two kimchi gates, restated from the reference for the evaluator tests, plus their closed forms.

  generic gate     kimchi/src/circuits/polynomials/generic.rs:83-120     2 constraints, degree-2, evaluated over d4
  poseidon gate    kimchi/src/circuits/polynomials/poseidon.rs:351-436   15 constraints, x^7 S-box cached with Store/Load,
                                                                          MDS constants, next-row cells; evaluated over d8
  complete add     kimchi/src/circuits/polynomials/complete_add.rs:103-222   7 constraints, three cached sub-expressions (x21, y21,
                                                                          x1^2), the literal 1, doublings; degree 3, over d4
  endomul          kimchi/src/circuits/polynomials/endosclmul.rs:475-549      12 constraints: the EndoCoefficient constant, squares (Pow 2),
                                                                          eight cached sub-expressions, three next-row cells; degree 4+
  endomul scalar   kimchi/src/circuits/polynomials/endomul_scalar.rs:174-220  11 constraints: Horner polynomials with the rational
                                                                          literals 11/6, -5/2, 2/3, eight cached c_func values, long
                                                                          double-and-add folds; degree 4
  combination      kimchi/src/circuits/argument.rs:201-214               selector * sum_k alpha^(e_k) * constraint_k

Column order used by both programs (what the test uploads):  0..14 witness (d8) | 15..29 coefficients8 (d8) | 30 selector.
The token builder is any object with the methods of proof_systems_b200.ExprProgram; `Recorder` below records the same program for
the oracle (opcodes 0 literal, 1 cell, 2 dup, 3 pow, 4 add, 5 mul, 6 sub, 7 store, 8 load)."""
import numpy as np

W, COEFF, SELECTOR = 0, 15, 30
SPONGE_WIDTH, PERM_SBOX = 3, 7                      # poseidon.rs:53, PlonkSpongeConstantsKimchi
STATE_ORDER = [0, 2, 3, 4, 1]                       # poseidon.rs:65-74
ROUND_EQUATIONS = [(0, (False, 1)), (1, (False, 2)), (2, (False, 3)), (3, (False, 4)), (4, (True, 0))]   # poseidon.rs:289-310


def round_to_cols(i):                               # poseidon.rs:76-80
    start = STATE_ORDER[i] * SPONGE_WIDTH
    return range(start, start + SPONGE_WIDTH)


class Recorder:
    """same interface as ExprProgram, keeps (ops, args, literals) for oracle.expr_eval"""
    def __init__(self):
        self.ops, self.args, self.literals, self.n_cached = [], [], [], 0
    def _t(self, o, a=0): self.ops.append(o); self.args.append(a); return self
    def literal(self, x): self.literals.append(np.asarray(x, dtype=np.uint64).reshape(4)); return self._t(0, len(self.literals) - 1)
    def cell(self, col, next_row=False): return self._t(1, col | (0x80000000 if next_row else 0))
    def dup(self): return self._t(2)
    def pow(self, n): return self._t(3, n)
    def add(self): return self._t(4)
    def mul(self): return self._t(5)
    def sub(self): return self._t(6)
    def store(self): self._t(7); self.n_cached += 1; return self.n_cached - 1
    def load(self, slot): return self._t(8, slot)


def generic_gate(p, alpha_pows):
    """selector * (alpha^e0 * c1 + alpha^e1 * c2),  c = l_coeff*l + r_coeff*r + o_coeff*o + m_coeff*l*r + constant  (generic.rs:88-117)"""
    p.cell(SELECTOR)
    for g in range(2):
        c, w = COEFF + 5 * g, W + 3 * g
        p.literal(alpha_pows[g])
        p.cell(c + 0).cell(w + 0).mul()
        p.cell(c + 1).cell(w + 1).mul().add()
        p.cell(c + 2).cell(w + 2).mul().add()
        p.cell(c + 3).cell(w + 0).mul().cell(w + 1).mul().add()
        p.cell(c + 4).add()
        p.mul()                      # alpha^e * constraint
        if g:
            p.add()
    p.mul()                          # index(gate_type) * combined_constraints (argument.rs:209-210)
    return p


def generic_closed_form(P, sel, w, coeff, alpha_pows, i, scale_w, scale_sel=1):
    """the same value from the formula, Python integers mod P; w, coeff: [15][len] canonical ints over d8; index i of d4"""
    j = scale_w * i
    tot = 0
    for g in range(2):
        l, r, o = w[3 * g][j], w[3 * g + 1][j], w[3 * g + 2][j]
        c = coeff[5 * g:5 * g + 5]
        tot += alpha_pows[g] * (c[0][j] * l + c[1][j] * r + c[2][j] * o + c[3][j] * l * r + c[4][j])
    return sel[scale_sel * i] * tot % P


def poseidon_gate(p, alpha_pows, mds):
    """selector * sum_k alpha^(e_k) * (w[target] - (rc + sum_c mds[j][c] * sboxed[c]))  (poseidon.rs:364-432);
    mds: [3][3] Montgomery limbs; the S-box powers are cached the first time they are used, as cache.cache(...) does"""
    p.cell(SELECTOR)
    idx = 0
    first = True
    for source, (target_next, target_round) in ROUND_EQUATIONS:
        slots = [None] * SPONGE_WIDTH
        for j, col in enumerate(round_to_cols(target_round)):
            p.literal(alpha_pows[idx])
            p.cell(W + col, target_next)                   # env.witness(target_row, col)
            p.cell(COEFF + idx)                            # rc
            for c, src in enumerate(round_to_cols(source)):
                p.literal(mds[j][c])
                if slots[c] is None:
                    p.cell(W + src).pow(PERM_SBOX)
                    slots[c] = p.store()
                else:
                    p.load(slots[c])
                p.mul().add()
            p.sub().mul()
            if not first:
                p.add()
            first = False
            idx += 1
    p.mul()
    return p


def poseidon_closed_form(P, sel, w, coeff, alpha_pows, mds, i, m, next_shift):
    tot, idx = 0, 0
    for source, (target_next, target_round) in ROUND_EQUATIONS:
        sboxed = [pow(w[c][i], PERM_SBOX, P) for c in round_to_cols(source)]
        for j, col in enumerate(round_to_cols(target_round)):
            tgt = w[col][(i + next_shift) % m] if target_next else w[col][i]
            tot += alpha_pows[idx] * (tgt - (coeff[idx][i] + sum(mds[j][c] * sboxed[c] for c in range(SPONGE_WIDTH))))
            idx += 1
    return sel[i] * tot % P



def complete_add_gate(p, alpha_pows, one):
    """selector * sum_k alpha^(e_k) * constraint_k for the 7 constraints of complete_add.rs:108-220; `one` is the Montgomery 1 (T::one());
    cache.cache(e) stores e the first time it is built and loads it afterwards, as Expr::Cache does in to_polish"""
    x1, y1, x2, y2, x3, y3, inf, same_x, sv, inf_z, x21_inv = range(W, W + 11)
    slots = {}

    def cached(name, build):
        if name in slots:
            p.load(slots[name])
        else:
            build()
            slots[name] = p.store()

    x21 = lambda: cached("x21", lambda: p.cell(x2).cell(x1).sub())
    y21 = lambda: cached("y21", lambda: p.cell(y2).cell(y1).sub())
    x1sq = lambda: cached("x1sq", lambda: p.cell(x1).cell(x1).mul())
    dbl = lambda: p.dup().add()                                     # Expr::Double
    cons = [
        lambda: (p.cell(x21_inv), x21(), p.mul(), p.literal(one).cell(same_x).sub(), p.sub()),          # zero_check: z_inv * z - (1 - r)
        lambda: (p.cell(same_x), x21(), p.mul()),                                                          #             r * z
        lambda: (p.cell(same_x), p.cell(sv), dbl(), p.cell(y1).mul(), x1sq(), dbl(), p.sub(), x1sq(), p.sub(), p.mul(),   # same_x * dbl_case
                 p.literal(one).cell(same_x).sub(), x21(), p.cell(sv).mul(), y21(), p.sub(), p.mul(), p.add()),           # + (1 - same_x) * add_case
        lambda: (p.cell(x1).cell(x2).add().cell(x3).add().cell(sv).cell(sv).mul().sub()),                  # x1 + x2 + x3 - s^2
        lambda: (p.cell(sv).cell(x1).cell(x3).sub().mul().cell(y1).sub().cell(y3).sub()),                  # s (x1 - x3) - y1 - y3
        lambda: (y21(), p.cell(same_x).cell(inf).sub(), p.mul()),                                          # y21 (same_x - inf)
        lambda: (y21(), p.cell(inf_z).mul().cell(inf).sub()),                                              # y21 inf_z - inf
    ]
    p.cell(SELECTOR)
    for k, build in enumerate(cons):
        p.literal(alpha_pows[k])
        build()
        p.mul()
        if k:
            p.add()
    p.mul()
    return p


def complete_add_closed_form(P, sel, w, alpha_pows, i, scale_w, scale_sel=1):
    j = scale_w * i
    x1, y1, x2, y2, x3, y3, inf, same_x, s, inf_z, x21_inv = (w[k][j] for k in range(11))
    x21, y21, x1sq = x2 - x1, y2 - y1, x1 * x1
    cons = [x21_inv * x21 - (1 - same_x), same_x * x21,
            same_x * (2 * s * y1 - 2 * x1sq - x1sq) + (1 - same_x) * (x21 * s - y21),
            x1 + x2 + x3 - s * s, s * (x1 - x3) - y1 - y3, y21 * (same_x - inf), y21 * inf_z - inf]
    return sel[scale_sel * i] * sum(a * c for a, c in zip(alpha_pows, cons)) % P



def endomul_scalar_literals(P):
    """the field constants of endomul_scalar.rs:188-197 as integers mod P"""
    inv = lambda a: pow(a, -1, P)
    c_coeffs = [0, 11 * inv(6) % P, (-5) * inv(2) % P, 2 * inv(3) % P]
    crumb_over_x = [(-6) % P, 11, (-6) % P, 1]
    d_minus_c = [(-1) % P, 3, (-1) % P]
    return c_coeffs, crumb_over_x, d_minus_c


def endomul_scalar_gate(p, alpha_pows, lit):
    """selector * sum_k alpha^(e_k) * constraint_k, the 11 constraints of endomul_scalar.rs:178-219.  `lit(v)` gives the Montgomery
    limbs of the integer v.  polynomial(coeffs, x) is the reference's Horner fold from T::zero() (endomul_scalar.rs:57-62)."""
    P = lit.modulus
    c_coeffs, crumb_over_x, d_minus_c = endomul_scalar_literals(P)
    n0, n8, a0, b0, a8, b8 = range(W, W + 6)
    xs = [W + 6 + i for i in range(8)]

    def polynomial(coeffs, x):
        p.literal(lit(0))                                   # T::zero()
        for c in reversed(coeffs):
            p.cell(x).mul().literal(lit(c)).add()           # acc * x + literal(c)

    slots = [None] * 8

    def c_func(i):                                           # cache.cache(polynomial(c_coeffs, xs[i]))
        if slots[i] is None:
            polynomial(c_coeffs, xs[i])
            slots[i] = p.store()
        else:
            p.load(slots[i])

    dbl = lambda: p.dup().add()

    def n8_constraint():
        p.cell(n0)
        for x in xs:
            dbl(); dbl(); p.cell(x).add()
        p.cell(n8).sub()

    def a8_constraint():
        p.cell(a0)
        for i in range(8):
            dbl(); c_func(i); p.add()
        p.cell(a8).sub()

    def b8_constraint():
        p.cell(b0)
        for i in range(8):
            dbl(); c_func(i); polynomial(d_minus_c, xs[i]); p.add(); p.add()     # acc.double() + (c_func + polynomial(d - c))
        p.cell(b8).sub()

    cons = [n8_constraint, a8_constraint, b8_constraint] + [(lambda x=x: (polynomial(crumb_over_x, x), p.cell(x).mul())) for x in xs]
    p.cell(SELECTOR)
    for k, build in enumerate(cons):
        p.literal(alpha_pows[k])
        build()
        p.mul()
        if k:
            p.add()
    p.mul()
    return p


def endomul_scalar_closed_form(P, sel, w, alpha_pows, i):
    c_coeffs, crumb_over_x, d_minus_c = endomul_scalar_literals(P)
    poly = lambda cs, x: sum(c * pow(x, k, P) for k, c in enumerate(cs)) % P
    n0, n8, a0, b0, a8, b8 = (w[k][i] for k in range(6))
    xs = [w[6 + k][i] for k in range(8)]
    n, a, b = n0, a0, b0
    for x in xs:
        n = (4 * n + x) % P
        a = (2 * a + poly(c_coeffs, x)) % P
        b = (2 * b + poly(c_coeffs, x) + poly(d_minus_c, x)) % P
    cons = [n - n8, a - a8, b - b8] + [poly(crumb_over_x, x) * x for x in xs]
    return sel[i] * sum(al * c for al, c in zip(alpha_pows, cons)) % P



def endomul_gate(p, alpha_pows, one, endo):
    """selector * sum_k alpha^(e_k) * constraint_k, the 12 constraints of endosclmul.rs:479-548; `one`, `endo` Montgomery limbs of 1 and
    of the endo coefficient (Constant(EndoCoefficient), resolved by the caller like every constant)"""
    xt, yt, inv, xp, yp, n, xr, yr, s1, s3, b1, b2, b3, b4 = W + 0, W + 1, W + 2, W + 4, W + 5, W + 6, W + 7, W + 8, W + 9, W + 10, W + 11, W + 12, W + 13, W + 14
    xs = lambda: p.cell(W + 4, True)            # env.witness_next(4)
    ys = lambda: p.cell(W + 5, True)
    slots = {}

    def cached(name, build):
        if name in slots:
            p.load(slots[name])
        else:
            build()
            slots[name] = p.store()

    dbl = lambda: p.dup().add()
    endo_minus_1 = lambda: p.literal(endo).literal(one).sub()
    xq = lambda name, b: cached(name, lambda: (p.literal(one).cell(b), endo_minus_1(), p.mul().add().cell(xt).mul()))     # (1 + b (endo - 1)) xt
    xq1, xq2 = (lambda: xq("xq1", b1)), (lambda: xq("xq2", b3))
    yq = lambda b: (p.cell(b), dbl(), p.literal(one).sub().cell(yt).mul())                                                # (2 b - 1) yt
    s1sq = lambda: cached("s1sq", lambda: p.cell(s1).pow(2))
    s3sq = lambda: cached("s3sq", lambda: p.cell(s3).pow(2))
    xp_xr = lambda: cached("xp_xr", lambda: p.cell(xp).cell(xr).sub())
    xr_xs = lambda: cached("xr_xs", lambda: (p.cell(xr), xs(), p.sub()))
    ys_yr = lambda: cached("ys_yr", lambda: (ys(), p.cell(yr).add()))
    yr_yp = lambda: cached("yr_yp", lambda: p.cell(yr).cell(yp).add())
    boolean = lambda b: (p.cell(b).pow(2).cell(b).sub())

    def n_constraint():
        p.cell(n); dbl(); p.cell(b1).add(); dbl(); p.cell(b2).add(); dbl(); p.cell(b3).add(); dbl(); p.cell(b4).add()
        p.cell(W + 6, True).sub()

    # the cache order of the reference: xq1, xq2, s1^2, s3^2, xp-xr, xr-xs, ys+yr, yr+yp are built BEFORE the constraint list
    # (endosclmul.rs:502-523); here each is built where it is first used — the value of every constraint is the same
    cons = [
        lambda: boolean(b1), lambda: boolean(b2), lambda: boolean(b3), lambda: boolean(b4),
        lambda: (xq1(), p.cell(xp).sub().cell(s1).mul(), yq(b2), p.cell(yp).sub(), p.sub()),
        lambda: (p.cell(xp), dbl(), s1sq(), p.sub(), xq1(), p.add(), xp_xr(), p.cell(s1).mul(), yr_yp(), p.add(), p.mul(),
                 p.cell(yp), dbl(), xp_xr(), p.mul(), p.sub()),
        lambda: (yr_yp(), p.pow(2), xp_xr(), p.pow(2), s1sq(), xq1(), p.sub().cell(xr).add().mul(), p.sub()),
        lambda: (xq2(), p.cell(xr).sub().cell(s3).mul(), yq(b4), p.cell(yr).sub(), p.sub()),
        lambda: (p.cell(xr), dbl(), s3sq(), p.sub(), xq2(), p.add(), xr_xs(), p.cell(s3).mul(), ys_yr(), p.add(), p.mul(),
                 p.cell(yr), dbl(), xr_xs(), p.mul(), p.sub()),
        lambda: (ys_yr(), p.pow(2), xr_xs(), p.pow(2), s3sq(), xq2(), p.sub(), xs(), p.add().mul(), p.sub()),
        n_constraint,
        lambda: (xp_xr(), xr_xs(), p.mul().cell(inv).mul().literal(one).sub()),
    ]
    p.cell(SELECTOR)
    for k, build in enumerate(cons):
        p.literal(alpha_pows[k])
        build()
        p.mul()
        if k:
            p.add()
    p.mul()
    return p


def endomul_closed_form(P, sel, w, alpha_pows, endo, i, m, next_shift):
    j = (i + next_shift) % m
    xt, yt, inv, xp, yp, n, xr, yr, s1, s3, b1, b2, b3, b4 = (w[k][i] for k in (0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14))
    xs, ys, n_next = w[4][j], w[5][j], w[6][j]
    xq1, xq2 = (1 + b1 * (endo - 1)) * xt, (1 + b3 * (endo - 1)) * xt
    yq1, yq2 = (2 * b2 - 1) * yt, (2 * b4 - 1) * yt
    cons = [b1 * b1 - b1, b2 * b2 - b2, b3 * b3 - b3, b4 * b4 - b4,
            (xq1 - xp) * s1 - (yq1 - yp),
            (2 * xp - s1 * s1 + xq1) * ((xp - xr) * s1 + yr + yp) - 2 * yp * (xp - xr),
            (yr + yp) ** 2 - (xp - xr) ** 2 * (s1 * s1 - xq1 + xr),
            (xq2 - xr) * s3 - (yq2 - yr),
            (2 * xr - s3 * s3 + xq2) * ((xr - xs) * s3 + ys + yr) - 2 * yr * (xr - xs),
            (ys + yr) ** 2 - (xr - xs) ** 2 * (s3 * s3 - xq2 + xs),
            16 * n + 8 * b1 + 4 * b2 + 2 * b3 + b4 - n_next,
            (xp - xr) * (xr - xs) * inv - 1]
    return sel[i] * sum(a * c for a, c in zip(alpha_pows, cons)) % P
