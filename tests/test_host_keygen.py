"""CPU: the host-side keygen pieces behind the C ABI (no device): permutation assembly, selector
compression and the `.pk` file layout, against the oracle's plain-Python restatements and structural invariants."""
import ctypes as C

import numpy as np
import pytest

from oracle import keygen as okg
from tests.util import pkg


def _random_copies(rng, ncols, n, m):
    cp = np.stack([rng.integers(0, ncols, m), rng.integers(0, n, m), rng.integers(0, ncols, m), rng.integers(0, n, m)], axis=1)
    # chains, repeats and self-copies on purpose
    cp[m // 2:m // 2 + 5] = cp[:5]
    cp[-1] = [cp[0, 0], cp[0, 1], cp[0, 0], cp[0, 1]]
    return cp.astype(np.uint32)


@pytest.mark.parametrize("k,ncols,m,seed", [(4, 3, 20, 1), (6, 5, 200, 2), (8, 7, 900, 3)])
def test_permutation_assembly_matches_oracle(k, ncols, m, seed):
    api = pkg().api
    n = 1 << k
    cp = _random_copies(np.random.default_rng(seed), ncols, n, m)
    got = api.keygen_permutation_mapping(k, ncols, cp)
    a = okg.Assembly(ncols, n)
    for lc, lr, rc, rr in cp.tolist():
        a.copy(lc, lr, rc, rr)
    assert got.tolist() == a.flat()
    # the product's own Python workload generator builds the same structure
    class Cs:
        pass
    cs = Cs(); cs.n = n; cs.permutation = [None] * ncols
    assert got.tolist() == pkg().synth.build_permutation_mapping(cs, cp.tolist()).tolist()
    # invariants: a permutation whose cycles are the connected components of the copy graph
    assert sorted(got.tolist()) == list(range(ncols * n))
    parent = list(range(ncols * n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for lc, lr, rc, rr in cp.tolist():
        parent[find(lc * n + lr)] = find(rc * n + rr)
    seen = set()
    for start in range(ncols * n):
        if start in seen:
            continue
        cyc, i = [], start
        while i not in seen:
            seen.add(i); cyc.append(i); i = int(got[i])
        assert len({find(c) for c in cyc}) == 1
        assert len(cyc) == sum(1 for c in range(ncols * n) if find(c) == find(start))


def test_permutation_assembly_bounds_failure():
    api = pkg().api
    with pytest.raises(pkg().ZkcError) as e:
        api.keygen_permutation_mapping(3, 2, [[0, 1, 2, 0]])     # column 2 of 2
    assert e.value.code == 12
    with pytest.raises(pkg().ZkcError) as e:
        api.keygen_permutation_mapping(3, 2, [[0, 8, 1, 0]])     # row 8 of 8
    assert e.value.code == 12


@pytest.mark.parametrize("seed", range(6))
def test_compress_selectors_matches_oracle(seed):
    api = pkg().api
    rng = np.random.default_rng(seed)
    k, S, max_degree = 5, int(rng.integers(1, 12)), int(rng.integers(3, 7))
    n = 1 << k
    # sparse, partly overlapping activations so that combinations form and exclusions matter
    act = (rng.random((S, n)) < rng.choice([0.05, 0.2, 0.5])).astype(np.uint8)
    md = rng.integers(0, max_degree + 1, S).astype(np.uint32)
    comb, root, clen, cols = api.compress_selectors(k, act, md, max_degree)
    assign, ocols = okg.compress_selectors(act.tolist(), md.tolist(), max_degree)
    assert [(int(a), int(b), int(c)) for a, b, c in zip(comb, root, clen)] == assign
    assert cols.tolist() == ocols
    # what the substitution needs: on every row each selector is "on" exactly where its column holds its root
    for s in range(S):
        assert ((cols[comb[s]] == root[s]) == act[s].astype(bool)).all()
    # degree bound of every combination: (largest member degree - 1) + members <= max_degree
    for c in range(cols.shape[0]):
        members = [s for s in range(S) if comb[s] == c]
        if all(md[s] > 0 for s in members):
            assert max(int(md[s]) for s in members) - 1 + len(members) <= max_degree


def test_compress_selectors_baseconfig_shape():
    """halo2-base's FlexGate selectors are all active on the same rows: nothing can be combined, each selector keeps a column
    whose values are its own 0/1 activations (what the synthetic circuits assume)"""
    api = pkg().api
    k, S = 6, 4
    act = np.zeros((S, 1 << k), dtype=np.uint8)
    act[:, ::4] = 1
    comb, root, clen, cols = api.compress_selectors(k, act, [4] * S, 4)
    assert comb.tolist() == list(range(S)) and root.tolist() == [1] * S and clen.tolist() == [1] * S
    assert (cols == act).all()


def test_pk_file_layout_and_headers():
    api = pkg().api
    lib = api.lib()
    k, ek, F, P, ns = 6, 8, 5, 6, 3
    L = api.pk_file_layout(k, ek, F, P, ns)
    n, en = 1 << k, 1 << ek
    want = 8 + 64 * (F + P) + ns * (n // 8) + 3 * (4 + 32 * en) + 2 * (4 + F * (4 + 32 * n)) + (4 + F * (4 + 32 * en)) \
        + 2 * (4 + P * (4 + 32 * n)) + (4 + P * (4 + 32 * en))
    assert L.total == want and L.fixed_commitments_off == 8 and L.selectors_off == 8 + 64 * (F + P)
    for be in (1, 0):
        buf = np.zeros(L.total, dtype=np.uint8)
        assert lib.zkc_pk_file_write_headers(api._hp(buf), C.c_size_t(buf.size), k, ek, F, P, ns, be) == 0
        assert bytes(buf[:4]) == (k.to_bytes(4, "big") if be else k.to_bytes(4, "little"))
        found = C.c_int(-1)
        assert lib.zkc_pk_file_check(api._hp(buf), C.c_size_t(buf.size), k, ek, F, P, ns, -1, C.byref(found)) == 0 and found.value == be
        assert lib.zkc_pk_file_check(api._hp(buf), C.c_size_t(buf.size), k, ek, F, P, ns, 1 - be, None) != 0
        assert lib.zkc_pk_file_check(api._hp(buf), C.c_size_t(buf.size - 1), k, ek, F, P, ns, -1, None) != 0
        bad = buf.copy(); bad[L.fixed_polys_off + 3 - (0 if be else 3)] ^= 1       # a count field
        assert lib.zkc_pk_file_check(api._hp(bad), C.c_size_t(bad.size), k, ek, F, P, ns, be, None) != 0
        assert lib.zkc_pk_file_check(api._hp(buf), C.c_size_t(buf.size), k, ek, F + 1, P, ns, -1, None) != 0
    r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    col = np.frombuffer((r - 1).to_bytes(32, "little") + r.to_bytes(32, "little"), dtype=np.uint64).reshape(2, 4)
    assert lib.zkc_fr_column_is_canonical(api._hp(col), C.c_size_t(1)) == 1 and lib.zkc_fr_column_is_canonical(api._hp(col), C.c_size_t(2)) == 0
