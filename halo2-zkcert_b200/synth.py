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
