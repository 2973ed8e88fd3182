"""Seeded, shape-faithful synthetic circuits for the BASELINE.json configs.

There is no Rust witness generator in this image, so the inputs `create_proof` receives — the
constraint system, fixed columns, copy constraints, advice columns and instances — are synthesised
with the shape of halo2-base's `BaseConfig` as halo2-zkcert configures it
(/root/reference/src/helpers.rs:97-172: `set_k(k)`, `set_lookup_bits(k-1)`, one instance column):

  * A "gate" advice columns, each with the FlexGate polynomial q * (a + b*c - d) over rotations
    0..3 and its own selector (a fixed column after selector compression);
  * one lookup-advice column range-checked against a fixed table 0..2^lookup_bits;
  * one fixed constants column, one instance column;
  * equality enabled on every advice column, the constants column and the instance column.

The witness satisfies every gate, lookup and copy constraint, so proofs verify.  Values are Python
ints (canonical); conversion to Montgomery limbs happens at the device boundary.
"""
import hashlib
import random

import numpy as np

from .circuit import ANY_ADVICE, ANY_FIXED, ANY_INSTANCE, R_MOD, ConstraintSystem

DELTA = 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2
ROOT_OF_UNITY = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c


def base_constraint_system(k, num_gate_cols):
    A = num_gate_cols
    advice_queries, fixed_queries = [], []
    for i in range(A):
        advice_queries += [(i, 0), (i, 1), (i, 2), (i, 3)]
    advice_queries.append((A, 0))                 # lookup advice
    CONST_COL, TABLE_COL = A, A + 1                # fixed: selectors 0..A-1, constants, table
    fixed_queries.append((CONST_COL, 0))           # from enable_equality
    for i in range(A):
        fixed_queries.append((i, 0))
    fixed_queries.append((TABLE_COL, 0))
    instance_queries = [(0, 0)]
    gates = []
    for i in range(A):
        a, b, c, d = (("advice", 4 * i + r) for r in range(4))
        q = ("fixed", 1 + i)
        gates.append([("product", q, ("sum", ("sum", a, ("product", b, c)), ("neg", d)))])
    lookups = [([("advice", 4 * A)], [("fixed", 1 + A)])]
    permutation = [(ANY_ADVICE, i) for i in range(A + 1)] + [(ANY_FIXED, CONST_COL), (ANY_INSTANCE, 0)]
    return ConstraintSystem(k, A + 1, A + 2, 1, advice_queries, fixed_queries, instance_queries, gates, lookups, permutation)


class SynthCircuit:
    """cs + fixed columns + copy constraints + a satisfying witness (all canonical Python ints)."""

    def __init__(self, cs, fixed, copies, advice, instances, name):
        self.cs, self.fixed, self.copies, self.advice, self.instances, self.name = cs, fixed, copies, advice, instances, name

    def transcript_repr(self):
        """Stand-in for vk.transcript_repr (upstream hashes the Debug string of the pinned vk, which
        cannot be reproduced without Rust): Blake2b-512 of the cs wire blob, reduced mod r."""
        h = hashlib.blake2b(self.cs.serialize(), digest_size=64, person=b"Halo2-Verify-Key").digest()
        return int.from_bytes(h, "little") % R_MOD


def _cell_value(rng, lookup_bits):
    x = rng.random()
    if x < 0.50:
        return rng.getrandbits(1)
    if x < 0.75:
        return rng.getrandbits(8)
    if x < 0.95:
        return rng.getrandbits(64)
    return rng.randrange(R_MOD)


def make_base_circuit(k, num_gate_cols, seed=0, fill=0.9, copy_frac=0.25, n_instances=32, name=None):
    cs = base_constraint_system(k, num_gate_cols)
    A, n = num_gate_cols, 1 << k
    U = cs.usable_rows()
    lookup_bits = k - 1
    rng = random.Random(seed)
    advice = [[0] * n for _ in range(A + 1)]
    fixed = [[0] * n for _ in range(A + 2)]
    free = []          # (perm column index, row) of cells no active gate constrains
    ngates = U // 4
    for i in range(A):
        col, sel = advice[i], fixed[i]
        for g in range(ngates):
            r0 = 4 * g
            if rng.random() < fill:
                a, b, c = _cell_value(rng, lookup_bits), _cell_value(rng, lookup_bits), _cell_value(rng, lookup_bits)
                col[r0], col[r0 + 1], col[r0 + 2], col[r0 + 3] = a, b, c, (a + b * c) % R_MOD
                sel[r0] = 1
            else:
                for r in range(r0, r0 + 4):
                    col[r] = _cell_value(rng, lookup_bits)
                    free.append((i, r))
    # lookup advice: range-checked cells; table 0..2^lookup_bits (rest zero)
    look = advice[A]
    for r in range(U):
        look[r] = rng.getrandbits(lookup_bits)
    table = fixed[A + 1]
    for r in range(min(U, 1 << lookup_bits)):
        table[r] = r
    # constants column
    consts = fixed[A]
    n_const = min(64, U)
    for r in range(n_const):
        consts[r] = r if r < 32 else rng.getrandbits(64)
    instances = [[rng.getrandbits(8) for _ in range(n_instances)]]   # 32 hash bytes (helpers.rs:167)
    # copy constraints: each destination is a distinct free cell; sources are never destinations
    rng.shuffle(free)
    want = min(len(free), int(copy_frac * n))
    dests = free[:want]
    dest_set = set(dests)
    P_LOOK, P_CONST, P_INST = A, A + 1, A + 2      # indices into cs.permutation
    copies = []
    for j, (dc, dr) in enumerate(dests):
        kind = j % 8
        if kind == 0 and n_const:
            sr = rng.randrange(n_const)
            val, src = consts[sr], (P_CONST, sr)
        elif kind == 1 and j // 8 < n_instances:
            sr = j // 8
            val, src = instances[0][sr], (P_INST, sr)
        elif kind in (2, 3):
            sr = rng.randrange(U)
            val, src = look[sr], (P_LOOK, sr)
        else:
            while True:
                sc, sr = rng.randrange(A), rng.randrange(4 * ngates)
                if (sc, sr) not in dest_set:
                    break
            val, src = advice[sc][sr], (sc, sr)
        advice[dc][dr] = val
        copies.append((src[0], src[1], dc, dr))
    return SynthCircuit(cs, fixed, copies, advice, instances, name or "base_k%d_a%d" % (k, A))


# ---- permutation keygen (halo2 `permutation::keygen::Assembly`; host logic, off the hot path) ----
def build_permutation_mapping(cs, copies):
    """Returns mapping as a flat int64 array: mapping[col * n + row] = col' * n + row'."""
    n, m = cs.n, len(cs.permutation)
    mapping = np.arange(m * n, dtype=np.int64)
    aux = np.arange(m * n, dtype=np.int64)
    sizes = np.ones(m * n, dtype=np.int64)
    mp, ax, sz = mapping.tolist(), aux.tolist(), sizes.tolist()
    for lc, lr, rc, rr in copies:
        left, right = lc * n + lr, rc * n + rr
        if ax[left] == ax[right]:
            continue
        if sz[ax[left]] < sz[ax[right]]:
            left, right = right, left
        lcyc, rcyc = ax[left], ax[right]
        sz[lcyc] += sz[rcyc]
        i = rcyc
        while True:
            ax[i] = lcyc
            i = mp[i]
            if i == rcyc:
                break
        mp[left], mp[right] = mp[right], mp[left]
    return np.array(mp, dtype=np.int64)


def sigma_values(cs, mapping):
    """sigma_col[row] = DELTA^col' * omega^row' as canonical ints (small circuits / tests)."""
    n, m = cs.n, len(cs.permutation)
    omega = pow(ROOT_OF_UNITY, 1 << (28 - cs.k), R_MOD)
    wp = [1] * n
    for i in range(1, n):
        wp[i] = wp[i - 1] * omega % R_MOD
    dp = [pow(DELTA, c, R_MOD) for c in range(m)]
    out = []
    for c in range(m):
        seg = mapping[c * n:(c + 1) * n]
        out.append([dp[int(v) // n] * wp[int(v) % n] % R_MOD for v in seg])
    return out


def ints_to_limbs(vals):
    """canonical ints -> (len, 4) uint64 little-endian limbs"""
    b = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(b, dtype=np.uint64).reshape(-1, 4).copy()


# ---- zkEVM-SHA256 "bit" circuit shape (BASELINE config 3) ---------------------------------------------
def sha_bit_constraint_system(k, num_bit_cols=112, num_word_cols=3, n_xor=None, n_maj=8):
    """~110 bit-valued advice columns constrained by many small gates, no lookups
    (/root/reference/src/sha256_bit_circuit.rs:52-54: Sha256CircuitConfig + one instance column;
    SURVEY §8 config 3).  Gates (all gated by one fixed selector q):
      booleanity  q * b * (1 - b)                       for every bit column
      xor         q * (a + b - 2ab - c(next row))       source columns -> derived column, rotation 1
      maj         q * (d - (ab + ac + bc - 2abc))       degree 4
      word        q * (w - sum_i 2^i b_i)               16 bits -> one word column
    """
    NB, NW = num_bit_cols, num_word_cols
    nS = NB // 2
    nD = NB - nS
    n_xor = nD - n_maj if n_xor is None else n_xor
    assert NB >= 16 * NW and n_xor + n_maj <= nD
    advice_queries = [(c, 0) for c in range(NB + NW)]
    rot1 = {}
    for j in range(n_xor):
        rot1[nS + j] = len(advice_queries)
        advice_queries.append((nS + j, 1))
    fixed_queries = [(0, 0), (1, 0)]         # q, constants
    instance_queries = [(0, 0)]
    q = ("fixed", 0)
    one = ("const", 1)
    A = lambda c: ("advice", c)
    mul = lambda x, y: ("product", x, y)
    add = lambda x, y: ("sum", x, y)
    neg = lambda x: ("neg", x)
    polys = []
    for c in range(NB):
        polys.append(mul(q, mul(A(c), add(one, neg(A(c))))))
    for j in range(n_xor):
        a, b = A((3 * j) % nS), A((3 * j + 1) % nS)
        cn = ("advice", rot1[nS + j])
        polys.append(mul(q, add(add(add(a, b), neg(("scaled", mul(a, b), 2))), neg(cn))))
    for j in range(n_maj):
        a, b, c = A((5 * j) % nS), A((5 * j + 2) % nS), A((5 * j + 4) % nS)
        d = A(nS + n_xor + j)
        ab, ac, bc = mul(a, b), mul(a, c), mul(b, c)
        polys.append(mul(q, add(d, neg(add(add(add(ab, ac), bc), neg(("scaled", mul(ab, c), 2)))))))
    for m in range(NW):
        acc = None
        for i in range(16):
            term = ("scaled", A(16 * m + i), 1 << i)
            acc = term if acc is None else add(acc, term)
        polys.append(mul(q, add(A(NB + m), neg(acc))))
    gates = [[p] for p in polys]
    permutation = [(ANY_INSTANCE, 0)] + [(ANY_ADVICE, NB + m) for m in range(min(NW, 2))] + [(ANY_FIXED, 1)]
    cs = ConstraintSystem(k, NB + NW, 2, 1, advice_queries, fixed_queries, instance_queries, gates, [], permutation)
    cs.sha_layout = dict(NB=NB, NW=NW, nS=nS, n_xor=n_xor, n_maj=n_maj)
    return cs


class LimbCircuit(SynthCircuit):
    """SynthCircuit whose columns are numpy canonical-limb arrays (n, 4) uint64 — for wide / tall shapes
    where Python-int lists would be too slow.  `.advice` / `.fixed` materialise int lists on demand."""

    def __init__(self, cs, fixed_limbs, copies, advice_limbs, instances, name):
        self.cs, self.fixed_limbs, self.copies, self.advice_limbs, self.instances, self.name = cs, fixed_limbs, copies, advice_limbs, instances, name

    @staticmethod
    def _ints(limbs):
        return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in limbs]

    @property
    def advice(self):
        return [self._ints(c) for c in self.advice_limbs]

    @property
    def fixed(self):
        return [self._ints(c) for c in self.fixed_limbs]


def make_sha_bit_circuit(k, num_bit_cols=112, num_word_cols=3, blocks=16, seed=0, name=None):
    """Witness for the SHA256-bit shape: `blocks` SHA blocks x 72 rows active (a 970-byte TBS like
    certs/example_cert_3.pem hashes in 16 blocks), zeros elsewhere."""
    cs = sha_bit_constraint_system(k, num_bit_cols, num_word_cols)
    lay = cs.sha_layout
    NB, NW, nS, n_xor, n_maj = lay["NB"], lay["NW"], lay["nS"], lay["n_xor"], lay["n_maj"]
    n = 1 << k
    U = cs.usable_rows()
    R = min(72 * blocks, U - 2)                 # active rows 0..R-1 (the xor gate reads row r+1)
    rng = np.random.default_rng(seed)
    bits = np.zeros((NB, n), dtype=np.uint64)
    bits[:nS, :R] = rng.integers(0, 2, size=(nS, R), dtype=np.uint64)
    bits[nS + n_xor + n_maj:, :R] = rng.integers(0, 2, size=(NB - nS - n_xor - n_maj, R), dtype=np.uint64)
    for j in range(n_xor):
        a, b = bits[(3 * j) % nS], bits[(3 * j + 1) % nS]
        bits[nS + j, 1:R + 1] = (a ^ b)[:R]
    for j in range(n_maj):
        a, b, c = bits[(5 * j) % nS], bits[(5 * j + 2) % nS], bits[(5 * j + 4) % nS]
        bits[nS + n_xor + j, :R] = ((a & b) | (a & c) | (b & c))[:R]
    words = np.zeros((NW, n), dtype=np.uint64)
    for m in range(NW):
        for i in range(16):
            words[m] += bits[16 * m + i] << np.uint64(i)
    def limbs(col):
        out = np.zeros((n, 4), dtype=np.uint64)
        out[:, 0] = col
        return out
    advice_limbs = [limbs(bits[c]) for c in range(NB)] + [limbs(words[m]) for m in range(NW)]
    q = np.zeros(n, dtype=np.uint64)
    q[:R] = 1
    consts = np.zeros(n, dtype=np.uint64)
    consts[:16] = np.arange(16, dtype=np.uint64)
    fixed_limbs = [limbs(q), limbs(consts)]
    # instances = two word cells (helpers.rs:255-258 exposes the digest as two field elements)
    rows = [5 % R, 7 % R]
    instances = [[int(words[0][rows[0]]), int(words[min(1, NW - 1)][rows[1]])]]
    copies = [(0, 0, 1, rows[0])]
    if NW >= 2:
        copies.append((0, 1, 2, rows[1]))
    else:
        instances = [[instances[0][0]]]
    # constants column cell 3 == a word cell forced to 3? keep it simple: tie constants[0] (=0) to an inactive word cell
    pconst = len(cs.permutation) - 1
    copies.append((pconst, 0, 1, U - 1))        # word column row U-1 is 0 (inactive)
    return LimbCircuit(cs, fixed_limbs, copies, advice_limbs, instances, name or "sha_bit_k%d_b%d" % (k, NB))


# ---- vectorised BaseConfig generator for tall circuits (aggregation k=22, BASELINE config 5) -----------
def _mul_add_u64(a, b, c):
    """(a + b*c) for uint64 arrays as three uint64 limbs (value < 2^129)"""
    m32 = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    b0, b1, c0, c1 = b & m32, b >> s32, c & m32, c >> s32
    p00, p01, p10, p11 = b0 * c0, b0 * c1, b1 * c0, b1 * c1
    mid = p01 + p10
    carry_mid = (mid < p01).astype(np.uint64)
    lo = p00 + (mid << s32)
    carry_lo = (lo < p00).astype(np.uint64)
    hi = p11 + (mid >> s32) + (carry_mid << s32) + carry_lo
    lo2 = lo + a
    c2 = (lo2 < lo).astype(np.uint64)
    hi2 = hi + c2
    top = (hi2 < hi).astype(np.uint64)
    return lo2, hi2, top


def _cells_u64(rng, shape):
    """50 % bits, 25 % bytes, 25 % 64-bit limbs (halo2-ecc style non-native limbs are range-checked words)"""
    kind = rng.integers(0, 4, size=shape, dtype=np.uint8)
    v = rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)
    v = np.where(kind < 2, v & np.uint64(1), np.where(kind == 2, v & np.uint64(0xFF), v))
    return v


def make_base_circuit_fast(k, num_gate_cols, seed=0, fill=0.9, copy_frac=0.25, n_instances=32, name=None):
    """Same shape and constraint semantics as make_base_circuit, generated with numpy (no uniform-Fr cells)."""
    cs = base_constraint_system(k, num_gate_cols)
    A, n = num_gate_cols, 1 << k
    U = cs.usable_rows()
    lookup_bits = k - 1
    rng = np.random.default_rng(seed)
    ngates = U // 4
    advice_limbs = [np.zeros((n, 4), dtype=np.uint64) for _ in range(A + 1)]
    fixed_limbs = [np.zeros((n, 4), dtype=np.uint64) for _ in range(A + 2)]
    active = rng.random((A, ngates)) < fill
    for i in range(A):
        cells = _cells_u64(rng, (ngates, 4))
        lo, hi, top = _mul_add_u64(cells[:, 0], cells[:, 1], cells[:, 2])
        act = active[i]
        col = advice_limbs[i]
        rows = np.arange(ngates) * 4
        for r in range(3):
            col[rows + r, 0] = cells[:, r]
        col[rows + 3, 0] = np.where(act, lo, cells[:, 3])
        col[rows + 3, 1] = np.where(act, hi, 0)
        col[rows + 3, 2] = np.where(act, top, 0)
        fixed_limbs[i][rows, 0] = act.astype(np.uint64)
    look = advice_limbs[A]
    look[:U, 0] = rng.integers(0, 1 << lookup_bits, size=U, dtype=np.uint64)
    tl = min(U, 1 << lookup_bits)
    fixed_limbs[A + 1][:tl, 0] = np.arange(tl, dtype=np.uint64)
    n_const = min(64, U)
    fixed_limbs[A][:n_const, 0] = np.where(np.arange(n_const) < 32, np.arange(n_const), rng.integers(0, 1 << 62, size=n_const)).astype(np.uint64)
    inst = rng.integers(0, 256, size=n_instances, dtype=np.uint64)
    instances = [[int(v) for v in inst]]
    # destinations: cells of inactive gates; sources: constants / instances / lookup cells / cells of ACTIVE gates
    fcol, fgate = np.nonzero(~active)
    order = rng.permutation(len(fcol) * 4)
    want = min(len(order), int(copy_frac * n))
    sel = order[:want]
    dcol, drow = fcol[sel // 4], fgate[sel // 4] * 4 + (sel % 4)
    kind = np.arange(want) % 8
    P_LOOK, P_CONST, P_INST = A, A + 1, A + 2
    scol = np.empty(want, dtype=np.int64)
    srow = np.empty(want, dtype=np.int64)
    acol, agate = np.nonzero(active)
    pick = rng.integers(0, len(acol), size=want)
    scol[:], srow[:] = acol[pick], agate[pick] * 4 + rng.integers(0, 4, size=want)
    m = kind == 0
    scol[m], srow[m] = P_CONST, rng.integers(0, n_const, size=int(m.sum()))
    m = (kind == 1) & (np.arange(want) // 8 < n_instances)
    scol[m], srow[m] = P_INST, (np.arange(want) // 8)[m]
    m = (kind == 2) | (kind == 3)
    scol[m], srow[m] = P_LOOK, rng.integers(0, U, size=int(m.sum()))
    inst_limbs = np.zeros((n_instances, 4), dtype=np.uint64)
    inst_limbs[:, 0] = inst
    sources = {**{i: advice_limbs[i] for i in range(A + 1)}, P_CONST: fixed_limbs[A], P_INST: inst_limbs}
    for sc in np.unique(scol):
        m = scol == sc
        vals = sources[int(sc)][srow[m]]
        for dc in np.unique(dcol[m]):
            mm = m & (dcol == dc)
            advice_limbs[int(dc)][drow[mm]] = sources[int(sc)][srow[mm]]
        del vals
    copies = np.stack([scol, srow, dcol.astype(np.int64), drow.astype(np.int64)], axis=1)
    return LimbCircuit(cs, fixed_limbs, copies.tolist() if want < (1 << 16) else copies, advice_limbs, instances,
                       name or "base_fast_k%d_a%d" % (k, A))


def build_permutation_mapping_fast(cs, copies):
    """Mapping for copy lists where every destination (columns 2,3) is a fresh singleton cell merged into its
    source's cycle — the structure make_base_circuit_fast emits.  Equivalent to build_permutation_mapping:
    merging a singleton `right` into the cycle of `left` swaps mapping[left] and mapping[right]."""
    n, m = cs.n, len(cs.permutation)
    mp = np.arange(m * n, dtype=np.int64)
    copies = np.asarray(copies, dtype=np.int64)
    left = copies[:, 0] * n + copies[:, 1]
    right = copies[:, 2] * n + copies[:, 3]
    # sequential semantics: process in order; group by `left` so repeated sources chain correctly
    order = np.argsort(left, kind="stable")
    left, right = left[order], right[order]
    # for a run of copies sharing the same left cell L with rights r1..rt (in order), the swaps give:
    #   mp[L] = rt, mp[rt] = r(t-1), ..., mp[r1] = old mp[L] = L
    start = np.r_[True, left[1:] != left[:-1]]
    end = np.r_[start[1:], True]
    prev = np.r_[0, right[:-1]]
    mp[right] = np.where(start, left, prev)
    mp[left[end]] = right[end]
    return mp


# ---- small shapes that exercise the generic paths (tests) --------------------------------------------------
def make_multi_lookup_circuit(k, seed=0, with_permutation=True):
    """Two lookups — one with two theta-compressed expressions per side (compressed values are arbitrary field
    elements, so the 256-bit sort and the permutation run on full-width keys), one single-column — plus a gate with
    negative rotations, a Scaled / Constant expression, two instance columns and (optionally) no permutation at all."""
    n = 1 << k
    rng = random.Random(seed)
    # advice: 0 = a, 1 = b (pair lookup), 2 = c (single lookup), 3 = d (gate: d = 3*a(prev) + c + 5)
    advice_queries = [(0, 0), (1, 0), (2, 0), (3, 0), (0, -1)]
    fixed_queries = [(0, 0), (1, 0), (2, 0)]        # t1, t2, q
    instance_queries = [(0, 0), (1, 0)] if with_permutation else []
    a, b, c, d, a_prev = (("advice", i) for i in range(5))
    t1, t2, q = (("fixed", i) for i in range(3))
    gate = ("product", q, ("sum", d, ("neg", ("sum", ("sum", ("scaled", a_prev, 3), c), ("const", 5)))))
    lookups = [([a, b], [t1, t2]), ([("product", q, c)], [t1])]
    permutation = [(ANY_ADVICE, 3), (ANY_INSTANCE, 0), (ANY_INSTANCE, 1)] if with_permutation else []
    cs = ConstraintSystem(k, 4, 3, 2 if with_permutation else 0, advice_queries, fixed_queries, instance_queries, [[gate]], lookups, permutation)
    U = cs.usable_rows()
    tsize = max(4, U // 2)
    fixed = [[0] * n for _ in range(3)]
    for i in range(tsize):
        fixed[0][i] = i
        fixed[1][i] = (i * i * 0x1234567 + 7 * pow(3, i, R_MOD)) % R_MOD
    # rows beyond the table repeat entry 0 so every usable table row is a valid pair
    for i in range(tsize, U):
        fixed[0][i], fixed[1][i] = fixed[0][0], fixed[1][0]
    advice = [[0] * n for _ in range(4)]
    for r in range(U):
        j = rng.randrange(tsize)
        advice[0][r], advice[1][r] = fixed[0][j], fixed[1][j]
        advice[2][r] = rng.randrange(tsize)
    for r in range(1, U):
        fixed[2][r] = rng.getrandbits(1)
        advice[3][r] = (3 * advice[0][r - 1] + advice[2][r] + 5) % R_MOD if fixed[2][r] else rng.randrange(R_MOD)
    # where q = 0 the second lookup's input is 0, which is in t1 (entry 0)
    copies, instances = [], []
    if with_permutation:
        rows = [r for r in range(1, U) if fixed[2][r]][:6]
        instances = [[advice[3][r] for r in rows[:4]], [advice[3][r] for r in rows[4:6]]]
        copies = [(1, i, 0, r) for i, r in enumerate(rows[:4])] + [(2, i, 0, r) for i, r in enumerate(rows[4:6])]
    return SynthCircuit(cs, fixed, copies, advice, instances, "multi_lookup_k%d" % k)
