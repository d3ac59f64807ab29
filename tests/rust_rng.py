"""Restatement (test infrastructure) of the random streams behind the reference's byte-exact serialization regressions
(poly-commitment/tests/commitment.rs:288-443): `o1_utils::tests::make_test_rng(Some(seed))` is rand 0.8.5's
`StdRng::from_seed` = rand_chacha 0.3 `ChaCha12Rng` (Cargo.lock:2622-2645), and field elements are drawn by ark-ff 0.5's
`UniformRand for Fp<MontBackend<_, 4>>`: four `next_u64` limbs (little-endian limb order), the unused top bit shaved, rejected
while >= the modulus, and the accepted limbs ARE the Montgomery representation.  Self-checking: the tests that use this
module reproduce the reference's hard-coded bytes, which they could not if any of this were restated wrongly."""
import struct

MASK32 = 0xFFFFFFFF


def _rotl(v, n):
    return ((v << n) & MASK32) | (v >> (32 - n))


def _quarter(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & MASK32; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & MASK32; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & MASK32; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & MASK32; s[b] = _rotl(s[b] ^ s[c], 7)


def chacha_block(key_words, counter, rounds=12):
    """One 64-byte block: 16 u32 words.  64-bit block counter in words 12-13, stream id 0 in words 14-15 (rand_chacha)."""
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [counter & MASK32, (counter >> 32) & MASK32, 0, 0]
    s = list(init)
    for _ in range(rounds // 2):
        _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13); _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
        _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12); _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
    return [(x + y) & MASK32 for x, y in zip(s, init)]


class StdRng:
    """rand 0.8.5 StdRng (ChaCha12, 64-word buffer = 4 blocks); only the u64 path is needed here."""

    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf, self.index = [], 64

    def _refill(self):
        self.buf = []
        for _ in range(4):
            self.buf += chacha_block(self.key, self.counter)
            self.counter += 1
        self.index = 0

    def next_u64(self) -> int:
        if self.index >= 64:
            self._refill()
        assert self.index % 2 == 0          # only u64 draws: the odd-index path of BlockRng::next_u64 never triggers
        lo, hi = self.buf[self.index], self.buf[self.index + 1]
        self.index += 2
        return (hi << 32) | lo

    def field_mont_limbs(self, modulus: int) -> list:
        """ark-ff 0.5 `Fp::rand`: the accepted raw limbs (the element's MONTGOMERY representation), little-endian u64 limbs."""
        while True:
            limbs = [self.next_u64() for _ in range(4)]
            limbs[3] &= (1 << 63) - 1       # num_bits_to_shave = 256 - 255
            v = sum(l << (64 * i) for i, l in enumerate(limbs))
            if v < modulus:
                return limbs
