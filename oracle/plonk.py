"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

CPU restatement of halo2-axiom `keygen_pk` (the parts create_proof reads) and `plonk::create_proof`
with ProverSHPLONK / ProverGWC and the Blake2b / Keccak256 transcripts (SURVEY.md §3.2, §8a rows
a7-a14, Appendix A.5-A.12).  Upstream: halo2_proofs 0.2.0 @4b42325 src/plonk/{prover,evaluation,
permutation,lookup,vanishing}.rs, src/poly/kzg/multiopen/{shplonk,gwc}, src/transcript.rs — not
vendored under /root/reference (Cargo.lock:1320-1336); reference call sites
/root/reference/src/helpers.rs:233,299 (gen_snark_shplonk -> create_proof).

Python drives the protocol (transcript via hashlib, RNG = the ChaCha restatement in
tests/pyref.py); every O(n) loop runs in the C++ oracle library through oracle/orc.py.  The
version-dependent behaviours (SURVEY §8c OPEN-1/2/3/5) are switches on `ProverOptions`.
"""
import hashlib

import numpy as np

from . import orc
from .orc import R_MOD, P_MOD

ANY_ADVICE, ANY_FIXED, ANY_INSTANCE = 0, 1, 2
DELTA = 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2


def M(v):
    """int -> (1, 4) Montgomery"""
    return orc.fr_from_ints([v])


def I(a):
    """(1, 4) Montgomery -> int"""
    return orc.fr_to_ints(a)[0]


class ProverOptions:
    def __init__(self, advice_blinding="axiom", blind_draws=False, random_poly="serial", point_format=0, zeta_choice=0,
                 lookup_fill="pse", random_poly_threads=1):
        self.advice_blinding = advice_blinding   # OPEN-1: "axiom" (last row := 1) | "pse" (last u rows random)
        self.blind_draws = blind_draws           # OPEN-2: per-commitment Blind(Fr::random) draws
        self.random_poly = random_poly           # OPEN-3: "serial" (n draws from the caller's rng) | "chunked" (one ChaCha20Rng per
        self.random_poly_threads = random_poly_threads   # rayon thread, seeded by rng.fill_bytes; thread-count dependent)
        self.lookup_fill = lookup_fill           # OPEN-9: "pse" (leftovers pop repeated rows from the end) | "axiom" (ascending rows)
        self.point_format = point_format         # OPEN-5: 0 = sign in bit 7, identity all-zero; 1 = sign bit 6, identity bit 7
        self.zeta_choice = zeta_choice           # OPEN-4 (does not change proof bytes)


# ---- transcripts (A.12) ---------------------------------------------------------------------------
def compress_point(pt, fmt=0):
    if pt is None:
        b = bytearray(32)
        if fmt == 1:
            b[31] |= 0x80
        return bytes(b)
    x, y = pt
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= ((y & 1) << 7) if fmt == 0 else ((y & 1) << 6)
    return bytes(b)


def _keccak_f(st):
    RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
          0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
          0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
          0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
    ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
    mask = (1 << 64) - 1
    rol = lambda v, r: ((v << r) | (v >> (64 - r))) & mask if r else v
    for rc in RC:
        C_ = [st[x][0] ^ st[x][1] ^ st[x][2] ^ st[x][3] ^ st[x][4] for x in range(5)]
        D = [C_[(x - 1) % 5] ^ rol(C_[(x + 1) % 5], 1) for x in range(5)]
        st = [[st[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = rol(st[x][y], ROT[x][y])
        st = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        st[0][0] ^= rc
    return st


def keccak256(data: bytes) -> bytes:
    """original Keccak-256 (pad 0x01), not SHA3-256"""
    rate = 136
    p = bytearray(data)
    p.append(0x01)
    while len(p) % rate:
        p.append(0)
    p[-1] |= 0x80
    st = [[0] * 5 for _ in range(5)]
    for off in range(0, len(p), rate):
        for i in range(rate // 8):
            st[i % 5][i // 5] ^= int.from_bytes(p[off + 8 * i: off + 8 * i + 8], "little")
        st = _keccak_f(st)
    out = b"".join(st[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


class EvmTranscriptWrite:
    """snark-verifier `EvmTranscript<G1Affine, NativeLoader, _, Vec<u8>>` (snark-verifier 0.1.6 @7011e8c
    src/system/halo2/transcript/evm.rs — un-vendored, /root/reference/Cargo.lock:2676-2693; instantiated by
    gen_evm_proof_shplonk, /root/reference/src/bin/cli.rs:519).  Recalled behaviour (SURVEY OPEN-7):
    the sponge buffer is a byte vector; points are absorbed / written as x || y, 32-byte big-endian each
    (uncompressed); scalars as 32-byte big-endian; squeeze = keccak256(buf ++ [0x01 if len(buf) == 32]),
    the digest replaces the buffer and is reduced mod r as a big-endian integer."""

    def __init__(self):
        self.buf = bytearray()
        self.proof = bytearray()

    def squeeze_challenge(self):
        data = bytes(self.buf) + (b"\x01" if len(self.buf) == 32 else b"")
        h = keccak256(data)
        self.buf = bytearray(h)
        return int.from_bytes(h, "big") % R_MOD

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.buf += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def common_scalar(self, s):
        self.buf += int(s).to_bytes(32, "big")

    def write_point(self, pt):
        self.common_point(pt)
        self.proof += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += int(s).to_bytes(32, "big")


class TranscriptWrite:
    """Blake2bWrite / Keccak256Write with Challenge255."""

    def __init__(self, kind="blake2b", point_format=0):
        self.kind = kind
        self.fmt = point_format
        self.proof = bytearray()
        if kind == "blake2b":
            self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        else:
            self.buf = bytearray()

    def _update(self, b):
        if self.kind == "blake2b":
            self.h.update(b)
        else:
            self.buf += b

    def squeeze_challenge(self):
        self._update(bytes([0]))
        if self.kind == "blake2b":
            d = self.h.copy().digest()
        else:
            d = keccak256(bytes(self.buf) + bytes([10])) + keccak256(bytes(self.buf) + bytes([11]))
        return int.from_bytes(d, "little") % R_MOD

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self._update(bytes([1]) + pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))

    def common_scalar(self, s):
        self._update(bytes([2]) + int(s).to_bytes(32, "little"))

    def write_point(self, pt):
        self.common_point(pt)
        self.proof += compress_point(pt, self.fmt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += int(s).to_bytes(32, "little")


# ---- expression evaluation over whole columns --------------------------------------------------------
def eval_expr_cols(e, get_query, nrows):
    t = e[0]
    if t == "const":
        return np.repeat(M(e[1]), nrows, axis=0)
    if t in ("advice", "fixed", "instance"):
        return get_query(t, e[1])
    if t == "neg":
        return orc.field_op("fr", "neg", eval_expr_cols(e[1], get_query, nrows))
    if t == "sum":
        return orc.vec_vec("add", eval_expr_cols(e[1], get_query, nrows), eval_expr_cols(e[2], get_query, nrows))
    if t == "product":
        return orc.vec_vec("mul", eval_expr_cols(e[1], get_query, nrows), eval_expr_cols(e[2], get_query, nrows))
    if t == "scaled":
        return orc.vec_scalar("mul", eval_expr_cols(e[1], get_query, nrows), M(e[2]))
    raise ValueError(t)


def eval_expr_scalar(e, get_query):
    t = e[0]
    if t == "const":
        return e[1] % R_MOD
    if t in ("advice", "fixed", "instance"):
        return get_query(t, e[1])
    if t == "neg":
        return (-eval_expr_scalar(e[1], get_query)) % R_MOD
    if t == "sum":
        return (eval_expr_scalar(e[1], get_query) + eval_expr_scalar(e[2], get_query)) % R_MOD
    if t == "product":
        return eval_expr_scalar(e[1], get_query) * eval_expr_scalar(e[2], get_query) % R_MOD
    if t == "scaled":
        return eval_expr_scalar(e[1], get_query) * e[2] % R_MOD
    raise ValueError(t)


# ---- keygen (what create_proof needs from the pk) -------------------------------------------------------
class ProvingKey:
    pass


def keygen(cs, fixed_mont, sigma_mont, g, g_lagrange, transcript_repr, zeta_choice=0):
    """fixed_mont / sigma_mont: lists of (n, 4) Montgomery Lagrange columns."""
    pk = ProvingKey()
    pk.cs = cs
    pk.k, pk.n = cs.k, cs.n
    pk.j = cs.degree()
    pk.dom = orc.domain_constants(pk.j, cs.k, zeta_choice)
    pk.zeta_choice = zeta_choice
    pk.ext_k = pk.dom["extended_k"]
    pk.ext_n = 1 << pk.ext_k
    pk.g, pk.g_lagrange = g, g_lagrange
    pk.transcript_repr = transcript_repr
    j, k = pk.j, pk.k
    pk.fixed_values = fixed_mont
    pk.fixed_polys = [orc.lagrange_to_coeff(j, k, f) for f in fixed_mont]
    pk.fixed_cosets = [orc.coeff_to_extended(j, k, p, zeta_choice) for p in pk.fixed_polys]
    pk.sigma_values = sigma_mont
    pk.sigma_polys = [orc.lagrange_to_coeff(j, k, f) for f in sigma_mont]
    pk.sigma_cosets = [orc.coeff_to_extended(j, k, p, zeta_choice) for p in pk.sigma_polys]
    bf = cs.blinding_factors()
    n = pk.n
    one = M(1)
    l0 = np.zeros((n, 4), dtype=np.uint64); l0[0] = one[0]
    l_last = np.zeros((n, 4), dtype=np.uint64); l_last[n - bf - 1] = one[0]
    l_blind = np.zeros((n, 4), dtype=np.uint64); l_blind[n - bf:] = one[0]
    ext = lambda lag: orc.coeff_to_extended(j, k, orc.lagrange_to_coeff(j, k, lag), zeta_choice)
    pk.l0, pk.l_last = ext(l0), ext(l_last)
    lb = ext(l_blind)
    ones = np.repeat(one, pk.ext_n, axis=0)
    pk.l_active_row = orc.vec_vec("sub", orc.vec_vec("sub", ones, pk.l_last), lb)
    pk.fixed_commitments = [commit(f, g_lagrange) for f in fixed_mont]
    pk.sigma_commitments = [commit(f, g_lagrange) for f in sigma_mont]
    return pk


def commit(poly, bases):
    """best_multiexp + to_affine -> (x, y) ints or None"""
    return orc.g1_to_ints(orc.best_multiexp(poly, bases[: poly.shape[0]]))[0]


def rotate_rows(a, rot, scale=1):
    return np.roll(a, -rot * scale, axis=0)


# ---- lookup permutation (A.6, PSE form) ---------------------------------------------------------------
def permute_expression_pair(cs, inp, tab, draw, fill="pse"):
    """fill = "pse": halo2 (PSE) permute_expression_pair — the leftover table values, ascending, are written to
    `repeated_input_rows.pop()`, i.e. to the repeated rows from the LAST one backwards.
    fill = "axiom": the rayon variant of the axiom fork as recalled (SURVEY OPEN-9) — first occurrences take their own value,
    every other row takes the next leftover table value (sorted table entries that are duplicates of their predecessor or
    absent from the input), ascending, in ascending row order.  Same multisets, different row assignment."""
    n, bf = cs.n, cs.blinding_factors()
    U = n - (bf + 1)
    a = orc.fr_to_ints(inp[:U])
    s = orc.fr_to_ints(tab[:U])
    a_sorted = sorted(a)
    left = {}
    for v in s:
        left[v] = left.get(v, 0) + 1
    s_perm = [None] * U
    repeated = []
    for row in range(U):
        v = a_sorted[row]
        if row == 0 or v != a_sorted[row - 1]:
            s_perm[row] = v
            c = left.get(v, 0)
            if c == 0:
                raise ValueError("ConstraintSystemFailure: lookup input not in table")
            left[v] = c - 1
        else:
            repeated.append(row)
    if fill == "axiom":
        repeated.reverse()        # pop() now yields the repeated rows in ascending order
    for v in sorted(left):
        for _ in range(left[v]):
            s_perm[repeated.pop()] = v
    assert not repeated
    a_tail = [draw() for _ in range(bf + 1)]
    s_tail = [draw() for _ in range(bf + 1)]
    return orc.fr_from_ints(a_sorted + a_tail), orc.fr_from_ints(s_perm + s_tail)


# ---- SHPLONK / GWC helpers ---------------------------------------------------------------------------
def lagrange_interpolate(points, evals):
    """coefficients (ints) of the polynomial of degree < len(points) through (points[i], evals[i])"""
    m = len(points)
    if m == 1:
        return [evals[0] % R_MOD]
    coeffs = [0] * m
    for j in range(m):
        # prod_{k != j} (X - x_k) / (x_j - x_k)
        num = [1]
        den = 1
        for kx in range(m):
            if kx == j:
                continue
            num = [(b - points[kx] * a) % R_MOD for a, b in zip(num + [0], [0] + num)]
            den = den * (points[j] - points[kx]) % R_MOD
        scale = evals[j] * pow(den, -1, R_MOD) % R_MOD
        for i in range(m):
            coeffs[i] = (coeffs[i] + num[i] * scale) % R_MOD
    return coeffs


def eval_ints(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


def vanishing_eval(roots, z):
    acc = 1
    for r in roots:
        acc = acc * (z - r) % R_MOD
    return acc


def construct_intermediate_sets(queries):
    """queries: list of (poly_id, point, eval).  Returns (rotation_sets, super_point_set) where
    rotation_sets = [(sorted points, [(poly_id, evals in point order)])] in first-appearance order."""
    super_points = sorted({q[1] for q in queries})
    poly_points = []          # [(poly_id, set(points))] in first-appearance order
    for pid, pt, _ in queries:
        for ent in poly_points:
            if ent[0] == pid:
                ent[1].add(pt)
                break
        else:
            poly_points.append((pid, {pt}))
    sets = []                 # [(frozenset points, [poly_id])]
    for pid, pts in poly_points:
        fs = frozenset(pts)
        for ent in sets:
            if ent[0] == fs:
                ent[1].append(pid)
                break
        else:
            sets.append((fs, [pid]))

    def get_eval(pid, pt):
        for q in queries:
            if q[0] == pid and q[1] == pt:
                return q[2]
        raise KeyError

    out = []
    for fs, pids in sets:
        pts = sorted(fs)
        out.append((pts, [(pid, [get_eval(pid, p) for p in pts]) for pid in pids]))
    return out, super_points


def _lincomb(polys_scalars, n):
    acc = np.zeros((n, 4), dtype=np.uint64)
    for poly, sc in polys_scalars:
        term = orc.vec_scalar("mul", poly, M(sc))
        if term.shape[0] < n:
            term = np.vstack([term, np.zeros((n - term.shape[0], 4), dtype=np.uint64)])
        acc = orc.vec_vec("add", acc, term)
    return acc


def _sub_low(poly, low_ints):
    out = poly.copy()
    m = len(low_ints)
    out[:m] = orc.vec_vec("sub", poly[:m], orc.fr_from_ints(low_ints))
    return out


def shplonk_prove(pk, transcript, polys, queries):
    """polys: {poly_id: (n,4) coefficient array}; queries: [(poly_id, point, eval)] in upstream order."""
    n = pk.n
    y = transcript.squeeze_challenge()
    rotation_sets, super_points = construct_intermediate_sets(queries)
    v = transcript.squeeze_challenge()
    # per set: r(X) for each poly, numerator combination, division by the set's vanishing polynomial
    ext_sets = []
    for pts, commitments in rotation_sets:
        ext_sets.append((pts, [(pid, evals, lagrange_interpolate(pts, evals)) for pid, evals in commitments]))
    h_x = np.zeros((n, 4), dtype=np.uint64)
    pv = 1
    for pts, commitments in ext_sets:
        n_x = np.zeros((n, 4), dtype=np.uint64)
        py = 1
        for pid, _, r_coeffs in commitments:
            numer = _sub_low(polys[pid], r_coeffs)
            n_x = orc.vec_vec("add", n_x, orc.vec_scalar("mul", numer, M(py)))
            py = py * y % R_MOD
        q = n_x
        for root in pts:
            q = orc.kate_division(q, M(root))
        q = np.vstack([q, np.zeros((n - q.shape[0], 4), dtype=np.uint64)])
        h_x = orc.vec_vec("add", h_x, orc.vec_scalar("mul", q, M(pv)))
        pv = pv * v % R_MOD
    transcript.write_point(commit(h_x, pk.g))
    u = transcript.squeeze_challenge()
    l_x = np.zeros((n, 4), dtype=np.uint64)
    z_diffs = []
    pv = 1
    for pts, commitments in ext_sets:
        diffs = [p for p in super_points if p not in pts]
        z_i = vanishing_eval(diffs, u)
        z_diffs.append(z_i)
        inner = np.zeros((n, 4), dtype=np.uint64)
        py = 1
        for pid, _, r_coeffs in commitments:
            r_eval = eval_ints(r_coeffs, u)
            contrib = _sub_low(polys[pid], [r_eval])
            inner = orc.vec_vec("add", inner, orc.vec_scalar("mul", contrib, M(py)))
            py = py * y % R_MOD
        l_x = orc.vec_vec("add", l_x, orc.vec_scalar("mul", inner, M(z_i * pv % R_MOD)))
        pv = pv * v % R_MOD
    zt_eval = vanishing_eval(super_points, u)
    l_x = orc.vec_vec("sub", l_x, orc.vec_scalar("mul", h_x, M(zt_eval)))
    assert I(orc.eval_poly(l_x, M(u))) == 0, "SHPLONK linearisation does not vanish at u"
    w = orc.kate_division(l_x, M(u))
    w = orc.vec_scalar("mul", w, M(pow(z_diffs[0], -1, R_MOD)))
    transcript.write_point(commit(w, pk.g))


def gwc_prove(pk, transcript, polys, queries):
    v = transcript.squeeze_challenge()
    points = []
    for _, pt, _ in queries:
        if pt not in points:
            points.append(pt)
    for z in points:
        acc = np.zeros((pk.n, 4), dtype=np.uint64)
        eval_acc = 0
        pv = 1      # powers(v): the first query at this point gets v^0
        for pid, pt, ev in queries:
            if pt != z:
                continue
            acc = orc.vec_vec("add", acc, orc.vec_scalar("mul", polys[pid], M(pv)))
            eval_acc = (eval_acc + ev * pv) % R_MOD
            pv = pv * v % R_MOD
        num = _sub_low(acc, [eval_acc])
        w = orc.kate_division(num, M(z))
        transcript.write_point(commit(w, pk.g))


def chunked_random_poly(rng, n, threads):
    """vanishing::Argument::commit of halo2 >= v2023_04 / the axiom fork as recalled (SURVEY OPEN-3): n_chunks =
    threads + (n % threads != 0) ChaCha20Rng instances, each seeded with 32 bytes of rng.fill_bytes (in order), fill
    consecutive chunks of n / threads coefficients with Fr::random.  Thread-count dependent by construction."""
    chunk = n // threads
    if chunk == 0:
        raise ValueError("more threads than coefficients")
    n_chunks = threads + (1 if n % threads else 0)
    if (n + chunk - 1) // chunk != n_chunks:
        raise ValueError("zip_eq length mismatch (upstream panics for this thread count)")
    seeds = [rng.fill_bytes(32) for _ in range(n_chunks)]
    parts = []
    for c, seed in enumerate(seeds):
        cnt = min(chunk, n - c * chunk)
        parts.append(orc.ChaCha20Rng(seed, 20).fr_random_bulk(cnt))
    return np.concatenate(parts)


# ---- create_proof (§3.2) -----------------------------------------------------------------------------
def create_proof(pk, advice_mont, instances, rng, transcript_kind="blake2b", multiopen="shplonk", opts=None, trace=None):
    """advice_mont: list of (n, 4) Montgomery columns as synthesised (rows >= usable are overwritten by
    the blinding policy); instances: [[int, ...]] per instance column; rng: object with fr_random().
    Returns proof bytes."""
    opts = opts or ProverOptions()
    cs = pk.cs
    n, k, j, bf = pk.n, pk.k, pk.j, cs.blinding_factors()
    U = n - (bf + 1)
    zc = pk.zeta_choice
    if transcript_kind == "evm":
        tr = EvmTranscriptWrite()
    elif transcript_kind == "poseidon":
        from .poseidon import PoseidonTranscriptWrite
        tr = PoseidonTranscriptWrite(opts.point_format)
    else:
        tr = TranscriptWrite(transcript_kind, opts.point_format)
    tr.common_scalar(pk.transcript_repr)
    draw = rng.fr_random
    rot_scale = 1 << (pk.ext_k - k)
    omega = I(pk.dom["omega"])

    # 1. instances
    inst_values = []
    for col in instances:
        if len(col) > U:
            raise ValueError("InstanceTooLarge")
        for v in col:
            tr.common_scalar(v)
        inst_values.append(orc.fr_from_ints(list(col) + [0] * (n - len(col))))
    inst_polys = [orc.lagrange_to_coeff(j, k, v) for v in inst_values]

    # 2. advice
    advice = [a.copy() for a in advice_mont]
    for a in advice:
        if opts.advice_blinding == "axiom":
            a[n - 1] = M(1)[0]
        else:
            a[U:] = orc.fr_from_ints([draw() for _ in range(bf + 1)])
    if opts.blind_draws:
        for _ in advice:
            draw()
    for a in advice:
        tr.write_point(commit(a, pk.g_lagrange))
    advice_polys = [orc.lagrange_to_coeff(j, k, a) for a in advice]

    # 3. theta
    theta = tr.squeeze_challenge()

    def lagrange_query(kind, qi):
        if kind == "advice":
            c, r = cs.advice_queries[qi]; return rotate_rows(advice[c], r)
        if kind == "fixed":
            c, r = cs.fixed_queries[qi]; return rotate_rows(pk.fixed_values[c], r)
        c, r = cs.instance_queries[qi]; return rotate_rows(inst_values[c], r)

    # 4. lookups: compress, permute, commit
    lookups = []
    for inp_exprs, tab_exprs in cs.lookups:
        def compress(exprs):
            acc = np.zeros((n, 4), dtype=np.uint64)
            for e in exprs:
                acc = orc.vec_vec("add", orc.vec_scalar("mul", acc, M(theta)), eval_expr_cols(e, lagrange_query, n))
            return acc
        comp_in, comp_tab = compress(inp_exprs), compress(tab_exprs)
        perm_in, perm_tab = permute_expression_pair(cs, comp_in, comp_tab, draw, opts.lookup_fill)
        if opts.blind_draws:
            draw()
        tr.write_point(commit(perm_in, pk.g_lagrange))
        if opts.blind_draws:
            draw()
        tr.write_point(commit(perm_tab, pk.g_lagrange))
        lookups.append(dict(comp_in=comp_in, comp_tab=comp_tab, perm_in=perm_in, perm_tab=perm_tab,
                            perm_in_poly=orc.lagrange_to_coeff(j, k, perm_in), perm_tab_poly=orc.lagrange_to_coeff(j, k, perm_tab)))

    # 5. beta, gamma
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()

    # 6. permutation grand products
    def perm_column_values(kind, idx):
        return {ANY_ADVICE: advice, ANY_FIXED: pk.fixed_values, ANY_INSTANCE: inst_values}[kind][idx]

    chunk = cs.permutation_chunk_len()
    perm_sets = []
    last_z = 1
    omega_pows = orc.powers(pk.dom["omega"], n)          # omega^i
    delta_omega = orc.vec_scalar("mul", omega_pows, M(beta))   # beta * omega^i, then * DELTA per column
    cols = cs.permutation
    for s0 in range(0, len(cols), chunk):
        sub = cols[s0:s0 + chunk]
        den = np.repeat(M(1), n, axis=0)
        for off, (kind, idx) in enumerate(sub):
            v = perm_column_values(kind, idx)
            t = orc.vec_vec("add", v, orc.vec_scalar("mul", pk.sigma_values[s0 + off], M(beta)))
            den = orc.vec_vec("mul", den, orc.vec_scalar("add", t, M(gamma)))
        ratio = orc.batch_invert(den)
        for off, (kind, idx) in enumerate(sub):
            v = perm_column_values(kind, idx)
            t = orc.vec_scalar("add", orc.vec_vec("add", v, delta_omega), M(gamma))
            ratio = orc.vec_vec("mul", ratio, t)
            delta_omega = orc.vec_scalar("mul", delta_omega, M(DELTA))
        z = orc.prefix_product(ratio, M(last_z))
        z[n - bf:] = orc.fr_from_ints([draw() for _ in range(bf)])
        last_z = I(z[U:U + 1])
        if opts.blind_draws:
            draw()
        tr.write_point(commit(z, pk.g_lagrange))
        perm_sets.append(dict(z=z, poly=orc.lagrange_to_coeff(j, k, z)))

    # 7. lookup grand products
    for lk in lookups:
        num = orc.vec_vec("mul", orc.vec_scalar("add", lk["comp_in"], M(beta)), orc.vec_scalar("add", lk["comp_tab"], M(gamma)))
        den = orc.vec_vec("mul", orc.vec_scalar("add", lk["perm_in"], M(beta)), orc.vec_scalar("add", lk["perm_tab"], M(gamma)))
        ratio = orc.vec_vec("mul", num, orc.batch_invert(den))
        z = orc.prefix_product(ratio, M(1))
        z[n - bf:] = orc.fr_from_ints([draw() for _ in range(bf)])
        if opts.blind_draws:
            draw()
        tr.write_point(commit(z, pk.g_lagrange))
        lk["z"] = z
        lk["z_poly"] = orc.lagrange_to_coeff(j, k, z)

    # 8. vanishing: random polynomial
    if opts.random_poly == "chunked":
        random_poly = chunked_random_poly(rng, n, opts.random_poly_threads)
    else:
        random_poly = rng.fr_random_bulk(n) if hasattr(rng, "fr_random_bulk") else orc.fr_from_ints([draw() for _ in range(n)])
    if opts.blind_draws:
        draw()
    tr.write_point(commit(random_poly, pk.g))

    # 9. y
    y = tr.squeeze_challenge()

    # 10. evaluate_h on the extended coset
    ext = lambda p: orc.coeff_to_extended(j, k, p, zc)
    advice_cosets = [ext(p) for p in advice_polys]
    inst_cosets = [ext(p) for p in inst_polys]

    def coset_query(kind, qi):
        if kind == "advice":
            c, r = cs.advice_queries[qi]; return rotate_rows(advice_cosets[c], r, rot_scale)
        if kind == "fixed":
            c, r = cs.fixed_queries[qi]; return rotate_rows(pk.fixed_cosets[c], r, rot_scale)
        c, r = cs.instance_queries[qi]; return rotate_rows(inst_cosets[c], r, rot_scale)

    en = pk.ext_n
    Y = M(y)
    fold = lambda value, term: orc.vec_vec("add", orc.vec_scalar("mul", value, Y), term)
    value = np.zeros((en, 4), dtype=np.uint64)
    for gate in cs.gates:
        for poly in gate:
            value = fold(value, eval_expr_cols(poly, coset_query, en))
    ones = np.repeat(M(1), en, axis=0)
    if perm_sets:
        zc_sets = [ext(s["poly"]) for s in perm_sets]
        last_rot = -(bf + 1)
        value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", ones, zc_sets[0]), pk.l0))
        zl = zc_sets[-1]
        value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", orc.vec_vec("mul", zl, zl), zl), pk.l_last))
        for si in range(1, len(zc_sets)):
            value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", zc_sets[si], rotate_rows(zc_sets[si - 1], last_rot, rot_scale)), pk.l0))
        zeta = I(pk.dom["g_coset"])
        cur_delta = orc.powers(pk.dom["extended_omega"], en, M(beta * zeta % R_MOD))
        coset_of = lambda kind, idx: {ANY_ADVICE: advice_cosets, ANY_FIXED: pk.fixed_cosets, ANY_INSTANCE: inst_cosets}[kind][idx]
        for si, s0 in enumerate(range(0, len(cols), chunk)):
            sub = cols[s0:s0 + chunk]
            left = rotate_rows(zc_sets[si], 1, rot_scale)
            for off, (kind, idx) in enumerate(sub):
                t = orc.vec_vec("add", coset_of(kind, idx), orc.vec_scalar("mul", pk.sigma_cosets[s0 + off], M(beta)))
                left = orc.vec_vec("mul", left, orc.vec_scalar("add", t, M(gamma)))
            right = zc_sets[si]
            for off, (kind, idx) in enumerate(sub):
                t = orc.vec_scalar("add", orc.vec_vec("add", coset_of(kind, idx), cur_delta), M(gamma))
                right = orc.vec_vec("mul", right, t)
                cur_delta = orc.vec_scalar("mul", cur_delta, M(DELTA))
            value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", left, right), pk.l_active_row))
    for (inp_exprs, tab_exprs), lk in zip(cs.lookups, lookups):
        zco, aco, sco = ext(lk["z_poly"]), ext(lk["perm_in_poly"]), ext(lk["perm_tab_poly"])

        def compress_coset(exprs):
            acc = np.zeros((en, 4), dtype=np.uint64)
            for e in exprs:
                acc = orc.vec_vec("add", orc.vec_scalar("mul", acc, M(theta)), eval_expr_cols(e, coset_query, en))
            return acc
        table_value = orc.vec_vec("mul", orc.vec_scalar("add", compress_coset(inp_exprs), M(beta)),
                                  orc.vec_scalar("add", compress_coset(tab_exprs), M(gamma)))
        a_minus_s = orc.vec_vec("sub", aco, sco)
        value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", ones, zco), pk.l0))
        value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", orc.vec_vec("mul", zco, zco), zco), pk.l_last))
        lhs = orc.vec_vec("mul", rotate_rows(zco, 1, rot_scale),
                          orc.vec_vec("mul", orc.vec_scalar("add", aco, M(beta)), orc.vec_scalar("add", sco, M(gamma))))
        value = fold(value, orc.vec_vec("mul", orc.vec_vec("sub", lhs, orc.vec_vec("mul", zco, table_value)), pk.l_active_row))
        value = fold(value, orc.vec_vec("mul", a_minus_s, pk.l0))
        value = fold(value, orc.vec_vec("mul", orc.vec_vec("mul", a_minus_s, orc.vec_vec("sub", aco, rotate_rows(aco, -1, rot_scale))),
                                        pk.l_active_row))

    # 11. h(X) = numerator / (X^n - 1); pieces; commitments
    h_ext = orc.divide_by_vanishing(j, k, value, zc)
    h_coeffs = orc.extended_to_coeff(j, k, h_ext, zc)
    q = j - 1
    if trace is not None:
        trace["h_coeffs"] = h_coeffs
    assert not h_coeffs[n * q:].any()
    h_pieces = [h_coeffs[i * n:(i + 1) * n] for i in range(q)]
    if opts.blind_draws:
        for _ in h_pieces:
            draw()
    for piece in h_pieces:
        tr.write_point(commit(piece, pk.g))

    # 12. x
    x = tr.squeeze_challenge()
    xn = pow(x, n, R_MOD)
    rot_point = lambda r: x * pow(omega, r, R_MOD) % R_MOD
    evalp = lambda poly, pt: I(orc.eval_poly(poly, M(pt)))

    # 13. evaluations
    advice_evals = [evalp(advice_polys[c], rot_point(r)) for c, r in cs.advice_queries]
    for e in advice_evals:
        tr.write_scalar(e)
    fixed_evals = [evalp(pk.fixed_polys[c], rot_point(r)) for c, r in cs.fixed_queries]
    for e in fixed_evals:
        tr.write_scalar(e)
    h_poly = np.zeros((n, 4), dtype=np.uint64)
    for piece in reversed(h_pieces):
        h_poly = orc.vec_vec("add", orc.vec_scalar("mul", h_poly, M(xn)), piece)
    random_eval = evalp(random_poly, x)
    tr.write_scalar(random_eval)
    sigma_evals = [evalp(p, x) for p in pk.sigma_polys]
    for e in sigma_evals:
        tr.write_scalar(e)
    x_next, x_last = rot_point(1), rot_point(-(bf + 1))
    perm_evals = []
    for si, s in enumerate(perm_sets):
        ev = dict(z=evalp(s["poly"], x), z_next=evalp(s["poly"], x_next))
        tr.write_scalar(ev["z"]); tr.write_scalar(ev["z_next"])
        if si != len(perm_sets) - 1:
            ev["z_last"] = evalp(s["poly"], x_last)
            tr.write_scalar(ev["z_last"])
        perm_evals.append(ev)
    x_inv = rot_point(-1)
    for lk in lookups:
        lk["evals"] = dict(z=evalp(lk["z_poly"], x), z_next=evalp(lk["z_poly"], x_next), a=evalp(lk["perm_in_poly"], x),
                           a_inv=evalp(lk["perm_in_poly"], x_inv), s=evalp(lk["perm_tab_poly"], x))
        for key in ("z", "z_next", "a", "a_inv", "s"):
            tr.write_scalar(lk["evals"][key])

    # 14. opening queries in upstream order (A.10)
    polys, queries = {}, []

    def add_query(pid, poly, pt, ev):
        polys[pid] = poly
        queries.append((pid, pt, ev))
    for (c, r), e in zip(cs.advice_queries, advice_evals):
        add_query(("advice", c), advice_polys[c], rot_point(r), e)
    for si, (s, ev) in enumerate(zip(perm_sets, perm_evals)):
        add_query(("perm_z", si), s["poly"], x, ev["z"])
        add_query(("perm_z", si), s["poly"], x_next, ev["z_next"])
    for si in reversed(range(len(perm_sets) - 1)):
        add_query(("perm_z", si), perm_sets[si]["poly"], x_last, perm_evals[si]["z_last"])
    for li, lk in enumerate(lookups):
        ev = lk["evals"]
        add_query(("lk_z", li), lk["z_poly"], x, ev["z"])
        add_query(("lk_a", li), lk["perm_in_poly"], x, ev["a"])
        add_query(("lk_s", li), lk["perm_tab_poly"], x, ev["s"])
        add_query(("lk_a", li), lk["perm_in_poly"], x_inv, ev["a_inv"])
        add_query(("lk_z", li), lk["z_poly"], x_next, ev["z_next"])
    for (c, r), e in zip(cs.fixed_queries, fixed_evals):
        add_query(("fixed", c), pk.fixed_polys[c], rot_point(r), e)
    for i, e in enumerate(sigma_evals):
        add_query(("sigma", i), pk.sigma_polys[i], x, e)
    add_query(("h",), h_poly, x, evalp(h_poly, x))
    add_query(("random",), random_poly, x, random_eval)

    if multiopen == "shplonk":
        shplonk_prove(pk, tr, polys, queries)
    else:
        gwc_prove(pk, tr, polys, queries)
    if trace is not None:
        trace.update(theta=theta, beta=beta, gamma=gamma, y=y, x=x)
    return bytes(tr.proof)
