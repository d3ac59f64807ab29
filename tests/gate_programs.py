"""RPN programs of two kimchi gates, restated from the reference for the evaluator tests, plus their closed forms.

  generic gate     kimchi/src/circuits/polynomials/generic.rs:83-120     2 constraints, degree-2, evaluated over d4
  poseidon gate    kimchi/src/circuits/polynomials/poseidon.rs:351-436   15 constraints, x^7 S-box cached with Store/Load,
                                                                          MDS constants, next-row cells; evaluated over d8
  complete add     kimchi/src/circuits/polynomials/complete_add.rs:103-222   7 constraints, three cached sub-expressions (x21, y21,
                                                                          x1^2), the literal 1, doublings; degree 3, over d4
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
