"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

CPU restatement of the host-side keygen pieces the product puts behind its C ABI: halo2's
`permutation::keygen::Assembly` and `ConstraintSystem::compress_selectors` (halo2_proofs 0.2.0 @4b42325
src/plonk/permutation/keygen.rs, src/plonk/circuit/compress_selectors.rs — not vendored under /root/reference,
Cargo.lock:1320-1336; reached from gen_pk, /root/reference/src/helpers.rs:213,265).  Plain Python, written
independently of the product's C++ (csrc/host/keygen.cpp) and of its Python workload generator.
"""


class Assembly:
    """permutation::keygen::Assembly: columns x n cells, `mapping` (successor in the cell's cycle), `aux` (cycle label),
    `sizes` (cycle length by label)."""

    def __init__(self, num_columns, n):
        self.n = n
        self.mapping = [[(c, r) for r in range(n)] for c in range(num_columns)]
        self.aux = [[(c, r) for r in range(n)] for c in range(num_columns)]
        self.sizes = [[1] * n for _ in range(num_columns)]

    def copy(self, lc, lr, rc, rr):
        if not (0 <= lr < self.n and 0 <= rr < self.n):
            raise ValueError("BoundsFailure")
        left, right = (lc, lr), (rc, rr)
        lcyc, rcyc = self.aux[lc][lr], self.aux[rc][rr]
        if lcyc == rcyc:
            return
        if self.sizes[lcyc[0]][lcyc[1]] < self.sizes[rcyc[0]][rcyc[1]]:
            left, right = right, left
            lcyc, rcyc = rcyc, lcyc
        self.sizes[lcyc[0]][lcyc[1]] += self.sizes[rcyc[0]][rcyc[1]]
        i = rcyc
        while True:
            self.aux[i[0]][i[1]] = lcyc
            i = self.mapping[i[0]][i[1]]
            if i == rcyc:
                break
        a, b = self.mapping[left[0]][left[1]], self.mapping[right[0]][right[1]]
        self.mapping[left[0]][left[1]], self.mapping[right[0]][right[1]] = b, a

    def flat(self):
        return [c2 * self.n + r2 for col in self.mapping for (c2, r2) in col]


def compress_selectors(activations, max_degrees, max_degree):
    """compress_selectors::process.  activations: list of 0/1 rows per selector.  Returns (assignments, columns):
    assignments[s] = (combination index, root j, combination length); columns[c][row] = root active there or 0."""
    S = len(activations)
    n = len(activations[0]) if S else 0
    columns, assign = [], [None] * S
    added = [False] * S

    def emit(members):
        col = [0] * n
        for j, s in enumerate(members, start=1):
            assign[s] = (len(columns), j, len(members))
            for r, a in enumerate(activations[s]):
                if a:
                    assert col[r] == 0, "selectors of one combination are disjoint"
                    col[r] = j
        columns.append(col)

    for s in range(S):
        assert max_degrees[s] <= max_degree
        if max_degrees[s] == 0:
            added[s] = True
            emit([s])
    excl = [[any(a and b for a, b in zip(activations[i], activations[j])) for j in range(i)] for i in range(S)]
    for i in range(S):
        if added[i]:
            continue
        added[i] = True
        d = max_degrees[i] - 1
        comb = [i]
        for j in range(i + 1, S):
            if d + len(comb) == max_degree:
                break
            if added[j] or any(excl[j][m] for m in comb):
                continue
            nd = max(d, max_degrees[j] - 1)
            if nd + len(comb) + 1 > max_degree:
                continue
            d = nd
            comb.append(j)
            added[j] = True
        emit(comb)
    return assign, columns
