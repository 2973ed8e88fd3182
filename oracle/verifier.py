"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

Restatement of halo2 `plonk::verify_proof` with VerifierSHPLONK / VerifierGWC over KZG on BN254
(halo2_proofs 0.2.0 @4b42325 src/plonk/verifier.rs, src/poly/kzg/multiopen/*/verifier.rs — not
vendored; Cargo.lock:1320-1336).  This is the "proof verifies" referee the reference's own tests use
(gen_snark_shplonk asserts verify_proof) — written independently of both provers:
plain Python integers, affine curve arithmetic, and either a real pairing check
(oracle/pairing.py) or, when the test knows the SRS trapdoor s, the equivalent check [s]W == P.
"""
from .plonk import (ANY_ADVICE, ANY_FIXED, ANY_INSTANCE, DELTA, construct_intermediate_sets, eval_expr_scalar, eval_ints,
                    keccak256, lagrange_interpolate, vanishing_eval)
from .orc import P_MOD, R_MOD
import hashlib

ROOT_OF_UNITY = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c


# ---- curve helpers (affine, Python ints) ---------------------------------------------------------------
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


def ec_mul(P, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = ec_add(acc, P)
        P = ec_add(P, P)
        k >>= 1
    return acc


def ec_neg(P):
    return None if P is None else (P[0], (-P[1]) % P_MOD)


def decompress_point(b, fmt=0):
    b = bytearray(b)
    if fmt == 0:
        sign = b[31] >> 7
        b[31] &= 0x7F
        if not any(b) and sign == 0:
            return None
    else:
        if b[31] & 0x80:
            return None
        sign = (b[31] >> 6) & 1
        b[31] &= 0x3F
    x = int.from_bytes(b, "little")
    if x >= P_MOD:
        raise ValueError("invalid point encoding")
    y2 = (x * x * x + 3) % P_MOD
    y = pow(y2, (P_MOD + 1) // 4, P_MOD)
    if y * y % P_MOD != y2:
        raise ValueError("point not on curve")
    if (y & 1) != sign:
        y = P_MOD - y
    return (x, y)


class TranscriptRead:
    def __init__(self, proof, kind="blake2b", point_format=0):
        self.kind, self.fmt, self.proof, self.pos = kind, point_format, bytes(proof), 0
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

    def common_scalar(self, s):
        self._update(bytes([2]) + int(s).to_bytes(32, "little"))

    def common_point(self, pt):
        self._update(bytes([1]) + pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))

    def read_point(self):
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


class EvmTranscriptRead:
    """reader side of oracle.plonk.EvmTranscriptWrite (64-byte big-endian uncompressed points)"""

    def __init__(self, proof):
        self.proof, self.pos, self.buf = bytes(proof), 0, bytearray()

    def squeeze_challenge(self):
        h = keccak256(bytes(self.buf) + (b"\x01" if len(self.buf) == 32 else b""))
        self.buf = bytearray(h)
        return int.from_bytes(h, "big") % R_MOD

    def common_scalar(self, s):
        self.buf += int(s).to_bytes(32, "big")

    def common_point(self, pt):
        self.buf += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")

    def read_point(self):
        x = int.from_bytes(self.proof[self.pos:self.pos + 32], "big")
        y = int.from_bytes(self.proof[self.pos + 32:self.pos + 64], "big")
        self.pos += 64
        if x >= P_MOD or y >= P_MOD or (y * y - x * x * x - 3) % P_MOD:
            raise ValueError("point not on curve")
        self.common_point((x, y))
        return (x, y)

    def read_scalar(self):
        s = int.from_bytes(self.proof[self.pos:self.pos + 32], "big")
        self.pos += 32
        if s >= R_MOD:
            raise ValueError("non-canonical scalar")
        self.common_scalar(s)
        return s


def l_i(x, n, omega, i):
    """Lagrange basis polynomial of row i (mod n) evaluated at x"""
    wi = pow(omega, i % n, R_MOD)
    return (pow(x, n, R_MOD) - 1) * pow(n, -1, R_MOD) % R_MOD * wi % R_MOD * pow(x - wi, -1, R_MOD) % R_MOD


class VerifyingKey:
    def __init__(self, cs, fixed_commitments, sigma_commitments, transcript_repr):
        self.cs, self.fixed_commitments, self.sigma_commitments, self.transcript_repr = cs, fixed_commitments, sigma_commitments, transcript_repr


def verify_proof(vk, g1_gen, instances, proof, check, transcript_kind="blake2b", multiopen="shplonk", point_format=0):
    """`check(left, right)` decides e(left, [s]G2) == e(right, G2); returns True/False."""
    cs = vk.cs
    n, k, bf = cs.n, cs.k, cs.blinding_factors()
    omega = pow(ROOT_OF_UNITY, 1 << (28 - k), R_MOD)
    if transcript_kind == "evm":
        tr = EvmTranscriptRead(proof)
    elif transcript_kind == "poseidon":
        from .poseidon import PoseidonTranscriptRead
        tr = PoseidonTranscriptRead(proof, point_format)
    else:
        tr = TranscriptRead(proof, transcript_kind, point_format)
    tr.common_scalar(vk.transcript_repr)
    for col in instances:
        for v in col:
            tr.common_scalar(v)
    advice_comms = [tr.read_point() for _ in range(cs.num_advice)]
    theta = tr.squeeze_challenge()
    lookup_perm = [(tr.read_point(), tr.read_point()) for _ in cs.lookups]
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    nsets = cs.num_permutation_sets()
    perm_comms = [tr.read_point() for _ in range(nsets)]
    lookup_z = [tr.read_point() for _ in cs.lookups]
    random_comm = tr.read_point()
    y = tr.squeeze_challenge()
    q = cs.degree() - 1
    h_comms = [tr.read_point() for _ in range(q)]
    x = tr.squeeze_challenge()
    xn = pow(x, n, R_MOD)
    rot_point = lambda r: x * pow(omega, r, R_MOD) % R_MOD

    # instance evaluations are computed by the verifier (KZG: QUERY_INSTANCE = false)
    instance_evals = []
    for c, r in cs.instance_queries:
        pt = rot_point(r)
        instance_evals.append(sum(v * l_i(pt, n, omega, i) for i, v in enumerate(instances[c])) % R_MOD)
    advice_evals = [tr.read_scalar() for _ in cs.advice_queries]
    fixed_evals = [tr.read_scalar() for _ in cs.fixed_queries]
    random_eval = tr.read_scalar()
    sigma_evals = [tr.read_scalar() for _ in cs.permutation]
    perm_evals = []
    for si in range(nsets):
        ev = dict(z=tr.read_scalar(), z_next=tr.read_scalar())
        if si != nsets - 1:
            ev["z_last"] = tr.read_scalar()
        perm_evals.append(ev)
    lookup_evals = []
    for _ in cs.lookups:
        lookup_evals.append(dict(z=tr.read_scalar(), z_next=tr.read_scalar(), a=tr.read_scalar(), a_inv=tr.read_scalar(), s=tr.read_scalar()))

    # expected h(x)
    l_last = l_i(x, n, omega, -(bf + 1))
    l_blind = sum(l_i(x, n, omega, -i) for i in range(1, bf + 1)) % R_MOD
    l_0 = l_i(x, n, omega, 0)
    active = (1 - (l_last + l_blind)) % R_MOD

    def get_query(kind, qi):
        return {"advice": advice_evals, "fixed": fixed_evals, "instance": instance_evals}[kind][qi]

    exprs = []
    for gate in cs.gates:
        for poly in gate:
            exprs.append(eval_expr_scalar(poly, get_query))
    if nsets:
        def col_eval(kind, idx):
            if kind == ANY_ADVICE:
                return advice_evals[cs.advice_queries.index((idx, 0))]
            if kind == ANY_FIXED:
                return fixed_evals[cs.fixed_queries.index((idx, 0))]
            return instance_evals[cs.instance_queries.index((idx, 0))]
        exprs.append(l_0 * (1 - perm_evals[0]["z"]) % R_MOD)
        zl = perm_evals[-1]["z"]
        exprs.append((zl * zl - zl) * l_last % R_MOD)
        for si in range(1, nsets):
            exprs.append((perm_evals[si]["z"] - perm_evals[si - 1]["z_last"]) * l_0 % R_MOD)
        chunk = cs.permutation_chunk_len()
        for si in range(nsets):
            cols = cs.permutation[si * chunk:(si + 1) * chunk]
            left = perm_evals[si]["z_next"]
            for off, (kind, idx) in enumerate(cols):
                left = left * (col_eval(kind, idx) + beta * sigma_evals[si * chunk + off] + gamma) % R_MOD
            right = perm_evals[si]["z"]
            cur = beta * x % R_MOD * pow(DELTA, si * chunk, R_MOD) % R_MOD
            for kind, idx in cols:
                right = right * (col_eval(kind, idx) + cur + gamma) % R_MOD
                cur = cur * DELTA % R_MOD
            exprs.append((left - right) * active % R_MOD)
    for (inp, tab), ev in zip(cs.lookups, lookup_evals):
        def compress(es):
            acc = 0
            for e in es:
                acc = (acc * theta + eval_expr_scalar(e, get_query)) % R_MOD
            return acc
        exprs.append(l_0 * (1 - ev["z"]) % R_MOD)
        exprs.append(l_last * (ev["z"] * ev["z"] - ev["z"]) % R_MOD)
        left = ev["z_next"] * (ev["a"] + beta) % R_MOD * (ev["s"] + gamma) % R_MOD
        right = ev["z"] * (compress(inp) + beta) % R_MOD * (compress(tab) + gamma) % R_MOD
        exprs.append((left - right) * active % R_MOD)
        exprs.append(l_0 * (ev["a"] - ev["s"]) % R_MOD)
        exprs.append((ev["a"] - ev["s"]) * (ev["a"] - ev["a_inv"]) % R_MOD * active % R_MOD)
    h_eval = 0
    for v in exprs:
        h_eval = (h_eval * y + v) % R_MOD
    expected_h_eval = h_eval * pow(xn - 1, -1, R_MOD) % R_MOD
    h_comm = None
    for c in reversed(h_comms):
        h_comm = ec_add(ec_mul(h_comm, xn), c)

    # queries (A.10)
    comms, queries = {}, []

    def add_query(pid, comm, pt, ev):
        comms[pid] = comm
        queries.append((pid, pt, ev))
    x_next, x_last, x_inv = rot_point(1), rot_point(-(bf + 1)), rot_point(-1)
    for (c, r), e in zip(cs.advice_queries, advice_evals):
        add_query(("advice", c), advice_comms[c], rot_point(r), e)
    for si in range(nsets):
        add_query(("perm_z", si), perm_comms[si], x, perm_evals[si]["z"])
        add_query(("perm_z", si), perm_comms[si], x_next, perm_evals[si]["z_next"])
    for si in reversed(range(nsets - 1)):
        add_query(("perm_z", si), perm_comms[si], x_last, perm_evals[si]["z_last"])
    for li, ev in enumerate(lookup_evals):
        add_query(("lk_z", li), lookup_z[li], x, ev["z"])
        add_query(("lk_a", li), lookup_perm[li][0], x, ev["a"])
        add_query(("lk_s", li), lookup_perm[li][1], x, ev["s"])
        add_query(("lk_a", li), lookup_perm[li][0], x_inv, ev["a_inv"])
        add_query(("lk_z", li), lookup_z[li], x_next, ev["z_next"])
    for (c, r), e in zip(cs.fixed_queries, fixed_evals):
        add_query(("fixed", c), vk.fixed_commitments[c], rot_point(r), e)
    for i, e in enumerate(sigma_evals):
        add_query(("sigma", i), vk.sigma_commitments[i], x, e)
    add_query(("h",), h_comm, x, expected_h_eval)
    add_query(("random",), random_comm, x, random_eval)

    if multiopen == "shplonk":
        yy = tr.squeeze_challenge()
        v = tr.squeeze_challenge()
        h1 = tr.read_point()
        u = tr.squeeze_challenge()
        h2 = tr.read_point()
        rotation_sets, super_points = construct_intermediate_sets(queries)
        outer, r_outer = None, 0
        z_0_diff_inv = z_0 = None
        pv = 1
        for i, (pts, commitments) in enumerate(rotation_sets):
            diffs = [p for p in super_points if p not in pts]
            z_diff_i = vanishing_eval(diffs, u)
            if i == 0:
                z_0 = vanishing_eval(pts, u)
                z_0_diff_inv = pow(z_diff_i, -1, R_MOD)
                z_diff_i = 1
            else:
                z_diff_i = z_diff_i * z_0_diff_inv % R_MOD
            inner, r_inner, py = None, 0, 1
            for pid, evals in commitments:
                r_x = lagrange_interpolate(pts, evals)
                r_inner = (r_inner + py * eval_ints(r_x, u)) % R_MOD
                inner = ec_add(inner, ec_mul(comms[pid], py))
                py = py * yy % R_MOD
            outer = ec_add(outer, ec_mul(inner, pv * z_diff_i % R_MOD))
            r_outer = (r_outer + pv * r_inner % R_MOD * z_diff_i) % R_MOD
            pv = pv * v % R_MOD
        outer = ec_add(outer, ec_mul(g1_gen, (-r_outer) % R_MOD))
        outer = ec_add(outer, ec_mul(h1, (-z_0) % R_MOD))
        outer = ec_add(outer, ec_mul(h2, u))
        ok = check(h2, outer)
    else:
        v = tr.squeeze_challenge()
        points = []
        for _, pt, _ in queries:
            if pt not in points:
                points.append(pt)
        ws = [tr.read_point() for _ in points]
        u = tr.squeeze_challenge()
        # sum_i u^i ( e(W_i, [s]G2) ) == sum_i u^i e(z_i W_i + C_i - [e_i]G, G2)
        left, right, pu = None, None, 1
        for z, w in zip(points, ws):
            cacc, eacc, pvv = None, 0, 1
            for pid, pt, ev in queries:
                if pt != z:
                    continue
                cacc = ec_add(cacc, ec_mul(comms[pid], pvv))
                eacc = (eacc + ev * pvv) % R_MOD
                pvv = pvv * v % R_MOD
            term = ec_add(ec_add(ec_mul(w, z), cacc), ec_mul(g1_gen, (-eacc) % R_MOD))
            left = ec_add(left, ec_mul(w, pu))
            right = ec_add(right, ec_mul(term, pu))
            pu = pu * u % R_MOD
        ok = check(left, right)
    return ok and tr.pos == len(proof)


def trapdoor_check(s):
    """the SRS secret is known (gen_srs is deterministic): e(L, [s]G2) == e(R, G2)  <=>  [s]L == R"""
    return lambda left, right: ec_mul(left, s) == right


def pairing_check(s_g2):
    from . import pairing

    def chk(left, right):
        return pairing.pairing_product_is_one([(left, s_g2), (ec_neg(right), pairing.G2_GEN)])
    return chk
