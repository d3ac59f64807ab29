"""The oracle's RPN evaluator (oracle/ntt_impl.h `expr_eval`, restating PolishToken::evaluate, kimchi/src/circuits/expr.rs:856-940)
pinned against the gates' closed forms computed with Python integers, and its failure modes against the reference's."""
import numpy as np
import pytest

import gate_programs as gp


def _setup(orc, fid, log_n, seed):
    n, m = 1 << log_n, 8 << log_n
    P = orc.FP_MODULUS if fid == orc.FP else orc.FQ_MODULUS
    rnd = lambda k, s: orc.to_mont(fid, orc.random_scalars(fid, k, seed=s))
    w = [rnd(m, seed + k) for k in range(15)]
    co = [rnd(m, seed + 100 + k) for k in range(15)]
    ints = lambda a: orc.limbs_to_ints(orc.from_mont(fid, a))
    return n, m, P, rnd, w, co, ints


@pytest.mark.parametrize("fid_name", ["FP", "FQ"])
def test_generic_gate_over_d4_matches_the_closed_form(orc, fid_name):
    fid = getattr(orc, fid_name)
    n, m, P, rnd, w, co, ints = _setup(orc, fid, 4, 1)
    sel4 = rnd(4 * n, 900)
    alphas = rnd(2, 901)
    rec = gp.generic_gate(gp.Recorder(), alphas)
    cols = [(a, 8) for a in w] + [(a, 8) for a in co] + [(sel4, 4)]
    got = ints(orc.expr_eval(fid, rec.ops, rec.args, rec.literals, cols, 4 * n))
    wi, ci, si, ai = [ints(a) for a in w], [ints(a) for a in co], ints(sel4), ints(alphas)
    for i in list(range(0, 4 * n, 7)) + [4 * n - 1]:
        assert got[i] == gp.generic_closed_form(P, si, wi, ci, ai, i, 2), i


@pytest.mark.parametrize("fid_name", ["FP", "FQ"])
def test_poseidon_gate_over_d8_matches_the_closed_form_and_accumulates(orc, fid_name):
    fid = getattr(orc, fid_name)
    n, m, P, rnd, w, co, ints = _setup(orc, fid, 3, 7)
    sel8, alphas, mds = rnd(m, 950), rnd(15, 951), rnd(9, 952).reshape(3, 3, 4)
    rec = gp.poseidon_gate(gp.Recorder(), alphas, mds)
    assert rec.n_cached == 15 and rec.ops.count(8) == 30            # 15 cached S-box powers, each reused twice
    cols = [(a, 8) for a in w] + [(a, 8) for a in co] + [(sel8, 8)]
    out = orc.expr_eval(fid, rec.ops, rec.args, rec.literals, cols, m)
    got = ints(out)
    wi, ci, si, ai = [ints(a) for a in w], [ints(a) for a in co], ints(sel8), ints(alphas)
    mi = [[ints(mds[r, c:c + 1])[0] for c in range(3)] for r in range(3)]
    for i in list(range(0, m, 5)) + [m - 8, m - 1]:                  # the last 8 rows wrap to row 0 (next = i + 8 mod m)
        assert got[i] == gp.poseidon_closed_form(P, si, wi, ci, ai, mi, i, m, 8), i
    # accumulate: t8 += eval
    acc = orc.expr_eval(fid, rec.ops, rec.args, rec.literals, cols, m, acc=out)
    assert ints(acc) == [2 * x % P for x in got]


@pytest.mark.parametrize("fid_name", ["FP", "FQ"])
def test_complete_add_gate_over_d4_matches_the_closed_form(orc, fid_name):
    fid = getattr(orc, fid_name)
    n, m, P, rnd, w, co, ints = _setup(orc, fid, 4, 21)
    sel4, alphas = rnd(4 * n, 960), rnd(7, 961)
    one = orc.to_mont(fid, orc.ints_to_limbs([1]))[0]
    rec = gp.complete_add_gate(gp.Recorder(), alphas, one)
    assert rec.n_cached == 3
    cols = [(a, 8) for a in w] + [(a, 8) for a in co] + [(sel4, 4)]
    got = ints(orc.expr_eval(fid, rec.ops, rec.args, rec.literals, cols, 4 * n))
    wi, si, ai = [ints(a) for a in w], ints(sel4), ints(alphas)
    for i in list(range(0, 4 * n, 5)) + [4 * n - 1]:
        assert got[i] == gp.complete_add_closed_form(P, si, wi, ai, i, 2), i


class _Lit:
    """integers -> Montgomery limbs of one field, for the gates that use rational literals"""
    def __init__(self, orc, fid, P):
        self.orc, self.fid, self.modulus = orc, fid, P
    def __call__(self, v):
        return self.orc.to_mont(self.fid, self.orc.ints_to_limbs([v % self.modulus]))[0]


@pytest.mark.parametrize("fid_name", ["FP", "FQ"])
def test_endomul_scalar_gate_matches_the_closed_form(orc, fid_name):
    fid = getattr(orc, fid_name)
    n, m, P, rnd, w, co, ints = _setup(orc, fid, 3, 33)
    sel8, alphas = rnd(m, 970), rnd(11, 971)
    rec = gp.endomul_scalar_gate(gp.Recorder(), alphas, _Lit(orc, fid, P))
    assert rec.n_cached == 8 and rec.ops.count(8) == 8           # eight c_func values, each reused once (in b8_expected)
    cols = [(a, 8) for a in w] + [(a, 8) for a in co] + [(sel8, 8)]
    got = ints(orc.expr_eval(fid, rec.ops, rec.args, rec.literals, cols, m))
    wi, si, ai = [ints(a) for a in w], ints(sel8), ints(alphas)
    for i in list(range(0, m, 3)) + [m - 1]:
        assert got[i] == gp.endomul_scalar_closed_form(P, si, wi, ai, i), i


@pytest.mark.parametrize("fid_name", ["FP", "FQ"])
def test_endomul_gate_matches_the_closed_form(orc, fid_name):
    fid = getattr(orc, fid_name)
    n, m, P, rnd, w, co, ints = _setup(orc, fid, 3, 44)
    sel8, alphas, endo = rnd(m, 980), rnd(12, 981), rnd(1, 982)[0]
    one = orc.to_mont(fid, orc.ints_to_limbs([1]))[0]
    rec = gp.endomul_gate(gp.Recorder(), alphas, one, endo)
    assert rec.n_cached == 8
    cols = [(a, 8) for a in w] + [(a, 8) for a in co] + [(sel8, 8)]
    got = ints(orc.expr_eval(fid, rec.ops, rec.args, rec.literals, cols, m))
    wi, si, ai, ei = [ints(a) for a in w], ints(sel8), ints(alphas), ints(endo[None])[0]
    for i in list(range(0, m, 3)) + [m - 8, m - 1]:
        assert got[i] == gp.endomul_closed_form(P, si, wi, ai, ei, i, m, 8), i


def test_failure_modes_of_the_reference(orc):
    one = orc.to_mont(orc.FP, orc.ints_to_limbs([1]))
    col = [(orc.to_mont(orc.FP, orc.random_scalars(orc.FP, 8, seed=3)), 8)]
    ok = orc.expr_eval(orc.FP, [0, 1, 4], [0, 0, 0], one, col, 8)              # 1 + cell
    assert ok.shape == (8, 4)
    for ops, args in ([[4], [0]],                       # Add on an empty stack: ExprError::EmptyStack
                      [[0, 0], [0, 0]],                 # two values left: assert_eq!(stack.len(), 1)
                      [[8], [0]],                       # Load before any Store
                      [[1], [5]],                       # column out of range
                      [[0], [3]]):                      # literal out of range
        with pytest.raises(ValueError):
            orc.expr_eval(orc.FP, ops, args, one, col, 8)
