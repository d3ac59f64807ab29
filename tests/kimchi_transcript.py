"""Restatement (test infrastructure, Python integers) of the Fiat-Shamir machinery `SRS::open` runs on the HOST in the reference —
it is outside the MSM/NTT hot path, and is restated here only so that the reference's byte-exact opening-proof
regression (poly-commitment/tests/commitment.rs:388-443) can be replayed around the device-resident rounds:
  Poseidon permutation / ArithmeticSponge      poseidon/src/permutation.rs:48-163, poseidon.rs:70-152 (PlonkSpongeConstantsKimchi,
                                               constants.rs:29-41: width 3, rate 2, 55 full rounds, x^7, full MDS, no initial ARK)
  DefaultFqSponge, ScalarChallenge::to_field   poseidon/src/sponge.rs
  endo coefficients                            poseidon/src/sponge.rs:43-47, poly-commitment/src/ipa.rs:214-231
  BW group map                                 groupmap/src/lib.rs
  ark-ff 0.5 Tonelli-Shanks square root        (which of the two roots comes out matters: it fixes y of u_base)
Checked against the reference's own Poseidon hash vectors (tests/golden/poseidon_kimchi.json) in tests/test_open_regression.py."""
import json
import os

FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
PARAMS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poseidon_kimchi.json")))


def sqrt_ark(a: int, p: int):
    """ark_ff::Field::sqrt for a field with generator 5 and two-adicity 32 (TonelliShanks precomputation): the root the
    algorithm lands on, None for a non-residue."""
    a %= p
    if a == 0:
        return 0
    t = (p - 1) >> 32
    z = pow(5, t, p)                       # QUADRATIC_NONRESIDUE_TO_TRACE = TWO_ADIC_ROOT_OF_UNITY
    w = pow(a, (t - 1) // 2, p)
    x = w * a % p
    b = x * w % p
    v = 32
    while b != 1:
        k, b2k = 0, b
        while b2k != 1:
            b2k = b2k * b2k % p
            k += 1
        if k == 32:
            return None
        j = v - k
        w = z
        for _ in range(1, j):
            w = w * w % p
        z = w * w % p
        b = b * z % p
        x = x * w % p
        v = k
    return x if x * x % p == a else None


class Poseidon:
    def __init__(self, field: str):
        self.p = FP if field == "fp" else FQ
        self.mds = PARAMS[field]["mds"]
        self.rc = PARAMS[field]["round_constants"]

    def permute(self, s):
        p = self.p
        for r in range(55):                # full_round: sbox, MDS, then the round constants (permutation.rs:55-70)
            s = [pow(x, 7, p) for x in s]
            s = [sum(m * x for m, x in zip(row, s)) % p for row in self.mds]
            s = [(x + c) % p for x, c in zip(s, self.rc[r])]
        return s


class ArithmeticSponge:
    """poseidon.rs:70-152"""

    def __init__(self, field: str):
        self.perm = Poseidon(field)
        self.p = self.perm.p
        self.state = [0, 0, 0]
        self.mode, self.n = "absorbed", 0

    def absorb(self, xs):
        for x in xs:
            if self.mode == "absorbed":
                if self.n == 2:
                    self.state = self.perm.permute(self.state)
                    self.n = 1
                    self.state[0] = (self.state[0] + x) % self.p
                else:
                    self.state[self.n] = (self.state[self.n] + x) % self.p
                    self.n += 1
            else:
                self.state[0] = (self.state[0] + x) % self.p
                self.mode, self.n = "absorbed", 1

    def squeeze(self) -> int:
        if self.mode == "squeezed":
            if self.n == 2:
                self.state = self.perm.permute(self.state)
                self.n = 1
                return self.state[0]
            self.n += 1
            return self.state[self.n - 1]
        self.state = self.perm.permute(self.state)
        self.mode, self.n = "squeezed", 1
        return self.state[0]


class DefaultFqSponge:
    """sponge.rs: DefaultFqSponge<P, PlonkSpongeConstantsKimchi>; `base` is the curve's base field ("fq" for Vesta)."""

    def __init__(self, base: str):
        self.sponge = ArithmeticSponge(base)
        self.base_p = self.sponge.p
        self.scalar_p = FP if base == "fq" else FQ
        self.last_squeezed = []

    def absorb_g(self, pts):               # pts: (x, y) canonical integers, or None for the identity
        self.last_squeezed = []
        for pt in pts:
            self.sponge.absorb([0] if pt is None else [pt[0]])
            self.sponge.absorb([0] if pt is None else [pt[1]])

    def absorb_fr(self, xs):
        self.last_squeezed = []
        for x in xs:
            if self.scalar_p < self.base_p:
                self.sponge.absorb([x])
            else:
                self.sponge.absorb([x >> 1])
                self.sponge.absorb([x & 1])

    def challenge_fq(self) -> int:         # squeeze_field
        self.last_squeezed = []
        return self.sponge.squeeze()

    def challenge(self) -> int:            # squeeze(CHALLENGE_LENGTH_IN_LIMBS = 2): a 128-bit scalar-field element
        if len(self.last_squeezed) < 2:
            x = self.sponge.squeeze()
            self.last_squeezed += [x & (2**64 - 1), (x >> 64) & (2**64 - 1)]       # HIGH_ENTROPY_LIMBS = 2
        lo, hi = self.last_squeezed[:2]
        self.last_squeezed = self.last_squeezed[2:]
        return lo | (hi << 64)


def endo_coefficient(p: int) -> int:
    """sponge.rs:43-47: GENERATOR^((p - 1) / 3), GENERATOR = 5"""
    return pow(5, (p - 1) // 3, p)


def scalar_challenge_to_field(pre: int, endo: int, p: int) -> int:
    """ScalarChallenge::to_field (sponge.rs:64-93), 128 bits"""
    a = b = 2
    for i in reversed(range(64)):
        a, b = 2 * a % p, 2 * b % p
        s = 1 if (pre >> (2 * i)) & 1 else p - 1
        if (pre >> (2 * i + 1)) & 1 == 0:
            b = (b + s) % p
        else:
            a = (a + s) % p
    return (a * endo + b) % p


class BWGroupMap:
    """groupmap/src/lib.rs for y^2 = x^3 + 5 over the field p"""

    def __init__(self, p: int, coeff_b: int = 5):
        self.p, self.B = p, coeff_b
        u = 1
        while self.curve_eqn(u) == 0:
            u += 1
        self.u, self.fu = u, self.curve_eqn(u)
        three_u2 = 3 * u * u % p
        self.inv_three_u_squared = pow(three_u2, -1, p)
        self.sqrt_neg_three_u_squared = sqrt_ark(-three_u2 % p, p)
        assert self.sqrt_neg_three_u_squared is not None
        self.sqrt_neg_three_u_squared_minus_u_over_2 = (self.sqrt_neg_three_u_squared - u) * pow(2, -1, p) % p

    def curve_eqn(self, x):
        return (x * x % self.p * x + self.B) % self.p

    def to_group(self, t: int):
        p = self.p
        t2 = t * t % p
        alpha_inv = (t2 + self.fu) * t2 % p
        alpha = pow(alpha_inv, -1, p) if alpha_inv else 0
        x1 = (self.sqrt_neg_three_u_squared_minus_u_over_2 - t2 * t2 % p * alpha % p * self.sqrt_neg_three_u_squared) % p
        x2 = (-self.u - x1) % p
        tf = (t2 + self.fu) % p
        x3 = (self.u - tf * tf % p * (alpha * tf % p) % p * self.inv_three_u_squared) % p
        for x in (x1, x2, x3):
            y = sqrt_ark(self.curve_eqn(x), p)
            if y is not None:
                return x, y
        raise AssertionError("get_xy")
