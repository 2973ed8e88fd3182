"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (one known answer: the first Grain round constant).

Restatement of the Poseidon sponge behind snark-verifier's `PoseidonTranscript<G1Affine, NativeLoader, Vec<u8>>`
as snark-verifier-sdk instantiates it for gen_snark_shplonk (POSEIDON_SPEC: T = 3, RATE = 2, R_F = 8, R_P = 57;
snark-verifier 0.1.6 @7011e8c src/util/hash/poseidon.rs + src/system/halo2/transcript/halo2.rs and the `poseidon`
crate's Grain parameter generation — all un-vendored: /root/reference/Cargo.lock:2676-2734; reference call sites
/root/reference/src/helpers.rs:233,299).  SURVEY.md OPEN-7.

Recalled behaviour:
  * round constants: Grain LFSR (80-bit state: field type 1 [2 bits], s-box 0 [4], field bits 254 [12], T [12], R_F [10],
    R_P [10], thirty ones; 160 warm-up bits; self-shrinking output), 254 bits MSB-first per element, rejection sampling;
    MDS = Cauchy matrix 1 / (x_i + y_j) with x, y the next 2T elements sampled WITHOUT rejection (reduced mod r);
  * permutation: R_F/2 full rounds, R_P partial rounds (s-box x^5 on element 0 only), R_F/2 full rounds, each round
    add-constants -> s-box -> MDS (the crate's optimised sparse form computes the same function);
  * sponge: state = [2^64, 0, 0]; absorb RATE elements into state[1..]; a chunk shorter than RATE gets a 1 right after
    it; a squeeze on an empty / exactly-full buffer runs one more permutation with only that 1; challenge = state[1];
    the state carries over (duplex);
  * transcript: scalars absorbed as is; points as (x mod r, y mod r); the proof stream holds 32-byte little-endian
    scalars and 32-byte compressed points, like halo2's own transcripts.
"""
from .orc import P_MOD, R_MOD

T, RATE, R_F, R_P = 3, 2, 8, 57


class Grain:
    def __init__(self, field_bits=254, t=T, r_f=R_F, r_p=R_P):
        bits = []

        def append(n, v):
            for i in reversed(range(n)):
                bits.append((v >> i) & 1)
        append(2, 1); append(4, 0); append(12, field_bits); append(12, t); append(10, r_f); append(10, r_p); append(30, (1 << 30) - 1)
        assert len(bits) == 80
        self.s = bits
        self.field_bits = field_bits
        for _ in range(160):
            self._new_bit()

    def _new_bit(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def next_bit(self):
        b = self._new_bit()
        while not b:
            self._new_bit()
            b = self._new_bit()
        return self._new_bit()

    def _take(self):
        v = 0
        for _ in range(self.field_bits):
            v = (v << 1) | self.next_bit()
        return v

    def next_field_element(self):
        while True:
            v = self._take()
            if v < R_MOD:
                return v

    def next_field_element_without_rejection(self):
        return self._take() % R_MOD


def generate_spec():
    g = Grain()
    constants = [[g.next_field_element() for _ in range(T)] for _ in range(R_F + R_P)]
    xs = [g.next_field_element_without_rejection() for _ in range(T)]
    ys = [g.next_field_element_without_rejection() for _ in range(T)]
    mds = [[pow(xs[i] + ys[j], -1, R_MOD) for j in range(T)] for i in range(T)]
    return constants, mds


_SPEC = None


def spec():
    global _SPEC
    if _SPEC is None:
        _SPEC = generate_spec()
    return _SPEC


def permute(state):
    constants, mds = spec()
    st = list(state)
    for r in range(R_F + R_P):
        st = [(a + c) % R_MOD for a, c in zip(st, constants[r])]
        if r < R_F // 2 or r >= R_F // 2 + R_P:
            st = [pow(a, 5, R_MOD) for a in st]
        else:
            st[0] = pow(st[0], 5, R_MOD)
        st = [sum(mds[i][j] * st[j] for j in range(T)) % R_MOD for i in range(T)]
    return st


class PoseidonSponge:
    def __init__(self):
        self.state = [1 << 64, 0, 0]
        self.buf = []

    def update(self, elements):
        self.buf.extend(int(e) % R_MOD for e in elements)

    def _absorb(self, chunk):
        for i, v in enumerate(chunk):
            self.state[1 + i] = (self.state[1 + i] + v) % R_MOD
        if len(chunk) < RATE:
            self.state[1 + len(chunk)] = (self.state[1 + len(chunk)] + 1) % R_MOD
        self.state = permute(self.state)

    def squeeze(self):
        buf, self.buf = self.buf, []
        exact = len(buf) % RATE == 0
        for i in range(0, len(buf), RATE):
            self._absorb(buf[i:i + RATE])
        if exact:
            self._absorb([])
        return self.state[1]


class PoseidonTranscriptWrite:
    def __init__(self, point_format=0):
        self.sponge = PoseidonSponge()
        self.proof = bytearray()
        self.fmt = point_format

    def squeeze_challenge(self):
        return self.sponge.squeeze()

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.sponge.update([pt[0] % R_MOD, pt[1] % R_MOD])

    def common_scalar(self, s):
        self.sponge.update([s])

    def write_point(self, pt):
        from .plonk import compress_point
        self.common_point(pt)
        self.proof += compress_point(pt, self.fmt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += int(s).to_bytes(32, "little")


class PoseidonTranscriptRead:
    def __init__(self, proof, point_format=0):
        self.sponge, self.proof, self.pos, self.fmt = PoseidonSponge(), bytes(proof), 0, point_format

    def squeeze_challenge(self):
        return self.sponge.squeeze()

    def common_scalar(self, s):
        self.sponge.update([s])

    def common_point(self, pt):
        self.sponge.update([pt[0] % R_MOD, pt[1] % R_MOD])

    def read_point(self):
        from .verifier import decompress_point
        pt = decompress_point(self.proof[self.pos:self.pos + 32], self.fmt)
        self.pos += 32
        if pt is None:
            raise ValueError("identity point in proof")
        self.common_point(pt)
        return pt

    def read_scalar(self):
        s = int.from_bytes(self.proof[self.pos:self.pos + 32], "little")
        self.pos += 32
        if s >= R_MOD:
            raise ValueError("non-canonical scalar")
        self.common_scalar(s)
        return s
