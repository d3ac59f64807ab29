"""RPN programs of two kimchi gates, restated from the reference for the evaluator tests, plus their closed forms.

  generic gate     kimchi/src/circuits/polynomials/generic.rs:83-120     2 constraints, degree-2, evaluated over d4
  poseidon gate    kimchi/src/circuits/polynomials/poseidon.rs:351-436   15 constraints, x^7 S-box cached with Store/Load,
                                                                          MDS constants, next-row cells; evaluated over d8
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
