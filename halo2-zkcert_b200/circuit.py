"""Host-side mirror of halo2_proofs `ConstraintSystem` (the part of `pk.vk.cs` create_proof reads) and
its wire format for the C ABI (`zkc_pk_*`, `zkc_prove`).

Upstream: halo2_proofs 0.2.0 @4b42325 src/plonk/circuit.rs (un-vendored; Cargo.lock:1320-1336);
SURVEY.md Appendix A.5.  A Rust host serialises `pk.vk.cs` into the same little-endian u32 stream
(INTEGRATION.md).  Nothing here is compute: column counts, query lists, gate expressions, lookup
expressions and the permutation column list.

Expressions are nested tuples:
  ("const", int) ("advice", query_index) ("fixed", query_index) ("instance", query_index)
  ("neg", e) ("sum", a, b) ("product", a, b) ("scaled", e, int)
"""
import struct

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001

# postfix opcodes of the serialised expression programs (also executed by the quotient kernel)
OP_CONST, OP_ADVICE, OP_FIXED, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE, OP_END = range(9)
ANY_ADVICE, ANY_FIXED, ANY_INSTANCE = 0, 1, 2


def expr_degree(e):
    t = e[0]
    if t == "const":
        return 0
    if t in ("advice", "fixed", "instance"):
        return 1
    if t in ("neg", "scaled"):
        return expr_degree(e[1])
    if t == "sum":
        return max(expr_degree(e[1]), expr_degree(e[2]))
    if t == "product":
        return expr_degree(e[1]) + expr_degree(e[2])
    raise ValueError(t)


class ConstraintSystem:
    def __init__(self, k, num_advice, num_fixed, num_instance, advice_queries, fixed_queries, instance_queries, gates, lookups,
                 permutation, minimum_degree=None):
        self.k = k
        self.n = 1 << k
        self.num_advice, self.num_fixed, self.num_instance = num_advice, num_fixed, num_instance
        self.advice_queries = list(advice_queries)      # [(column, rotation)]
        self.fixed_queries = list(fixed_queries)
        self.instance_queries = list(instance_queries)
        self.gates = [list(g) for g in gates]            # [[polynomial expression, ...], ...] in cs.gates order
        self.lookups = [(list(i), list(t)) for i, t in lookups]   # [(input_expressions, table_expressions)]
        self.permutation = list(permutation)             # [(ANY_*, column index)] in enable_equality order
        self.minimum_degree = minimum_degree

    # ---- A.5 numbers ---------------------------------------------------------------------------
    def blinding_factors(self):
        per_col = [0] * self.num_advice
        for c, _ in self.advice_queries:
            per_col[c] += 1
        factors = max([3] + per_col)
        return factors + 2

    def degree(self):
        d = 3  # permutation::Argument::required_degree()
        for inp, tab in self.lookups:
            di = max([1] + [expr_degree(e) for e in inp])
            dt = max([1] + [expr_degree(e) for e in tab])
            d = max(d, max(4, 2 + di + dt))
        for g in self.gates:
            for p in g:
                d = max(d, expr_degree(p))
        return max(d, self.minimum_degree or 1)

    def usable_rows(self):
        return self.n - (self.blinding_factors() + 1)

    def permutation_chunk_len(self):
        return self.degree() - 2

    def num_permutation_sets(self):
        c = self.permutation_chunk_len()
        return (len(self.permutation) + c - 1) // c

    def query_index(self, kind, column, rotation):
        qs = {ANY_ADVICE: self.advice_queries, ANY_FIXED: self.fixed_queries, ANY_INSTANCE: self.instance_queries}[kind]
        return qs.index((column, rotation))

    # ---- wire format ---------------------------------------------------------------------------
    def _emit_expr(self, e, words, consts):
        t = e[0]
        if t == "const":
            words += [OP_CONST, len(consts)]
            consts.append(e[1] % R_MOD)
        elif t == "advice":
            words += [OP_ADVICE, e[1]]
        elif t == "fixed":
            words += [OP_FIXED, e[1]]
        elif t == "instance":
            words += [OP_INSTANCE, e[1]]
        elif t == "neg":
            self._emit_expr(e[1], words, consts)
            words += [OP_NEG, 0]
        elif t == "sum":
            self._emit_expr(e[1], words, consts)
            self._emit_expr(e[2], words, consts)
            words += [OP_ADD, 0]
        elif t == "product":
            self._emit_expr(e[1], words, consts)
            self._emit_expr(e[2], words, consts)
            words += [OP_MUL, 0]
        elif t == "scaled":
            self._emit_expr(e[1], words, consts)
            words += [OP_SCALE, len(consts)]
            consts.append(e[2] % R_MOD)
        else:
            raise ValueError(t)

    def program(self, exprs):
        """postfix program for a list of expressions; each ends with OP_END"""
        words, consts = [], []
        for e in exprs:
            self._emit_expr(e, words, consts)
            words += [OP_END, 0]
        return words, consts

    def serialize(self):
        """little-endian blob:
        magic 'ZKCS', version, k, num_advice, num_fixed, num_instance, minimum_degree(0 = none),
        nq_advice, (col, rot as i32)*, nq_fixed, ..., nq_instance, ...,
        n_perm, (kind, col)*,
        gate program: n_words, words*, n_consts, consts (32 B canonical LE)*,
        n_lookups, per lookup: n_input_exprs, input program, n_table_exprs, table program
        """
        out = [b"ZKCS", struct.pack("<6I", 1, self.k, self.num_advice, self.num_fixed, self.num_instance, self.minimum_degree or 0)]
        for qs in (self.advice_queries, self.fixed_queries, self.instance_queries):
            out.append(struct.pack("<I", len(qs)))
            for c, r in qs:
                out.append(struct.pack("<Ii", c, r))
        out.append(struct.pack("<I", len(self.permutation)))
        for kind, c in self.permutation:
            out.append(struct.pack("<II", kind, c))

        def prog(exprs):
            w, cs = self.program(exprs)
            b = [struct.pack("<I", len(w) // 2)]
            b.append(struct.pack("<%dI" % len(w), *w))
            b.append(struct.pack("<I", len(cs)))
            for c in cs:
                b.append(int(c).to_bytes(32, "little"))
            return b"".join(b)

        polys = [p for g in self.gates for p in g]
        out.append(struct.pack("<I", len(polys)))
        out.append(prog(polys))
        out.append(struct.pack("<I", len(self.lookups)))
        for inp, tab in self.lookups:
            out.append(struct.pack("<I", len(inp)))
            out.append(prog(inp))
            out.append(struct.pack("<I", len(tab)))
            out.append(prog(tab))
        return b"".join(out)
