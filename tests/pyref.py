"""Python big-integer ground truth used to pin the oracle (tests only).

Everything here is the textbook definition: modular arithmetic with Python ints, affine
short-Weierstrass addition, O(n^2) DFT.  It shares no code with oracle/ or the product.
"""
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
ROOT_OF_UNITY = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
DELTA = 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2
S = 28
G1_GEN = (1, 2)


def ec_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return (x3, (lam * (x1 - x3) - y1) % P_MOD)


def ec_neg(P):
    return None if P is None else (P[0], (-P[1]) % P_MOD)


def ec_mul(P, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = ec_add(acc, P)
        P = ec_add(P, P)
        k >>= 1
    return acc


def ec_msm(scalars, points):
    acc = None
    for s, p in zip(scalars, points):
        acc = ec_add(acc, ec_mul(p, s))
    return acc


def omega_for(k):
    return pow(ROOT_OF_UNITY, 1 << (S - k), R_MOD)


def dft(a, omega):
    n = len(a)
    return [sum(a[i] * pow(omega, i * j, R_MOD) for i in range(n)) % R_MOD for j in range(n)]


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


# ---- ChaCha (RFC 7539 core, 64-bit counter variant used by rand_chacha) -------------------
def _rotl(v, c):
    return ((v << c) & 0xFFFFFFFF) | (v >> (32 - c))


def chacha_block(key_words, counter, rounds):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [counter & 0xFFFFFFFF, counter >> 32, 0, 0]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class ChaChaRng:
    """rand_chacha ChaCha{8,12,20}Rng restricted to next_u64 / next_u32."""

    def __init__(self, seed32: bytes, rounds=20):
        self.key = [int.from_bytes(seed32[4 * i:4 * i + 4], "little") for i in range(8)]
        self.rounds = rounds
        self.counter = 0
        self.buf = []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, self.rounds)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def fr_random(self):
        v = 0
        for i in range(8):
            v |= self.next_u64() << (64 * i)
        return v % R_MOD

    def fill_bytes(self, nbytes):
        """rand_core BlockRng::fill_bytes for a multiple of 4 bytes: whole keystream words, little-endian"""
        assert nbytes % 4 == 0
        return b"".join(self.next_u32().to_bytes(4, "little") for _ in range(nbytes // 4))


def seed_from_u64(state):
    """rand_core SeedableRng::seed_from_u64: PCG32 expansion into a 32-byte seed."""
    MUL = 6364136223846793005
    INC = 11634580027462260723
    out = b""
    for _ in range(8):
        state = (state * MUL + INC) & 0xFFFFFFFFFFFFFFFF
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF
        out += x.to_bytes(4, "little")
    return out
