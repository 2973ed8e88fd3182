"""GPU parity: best_multiexp / ParamsKZG through the C ABI vs the CPU oracle (bit-exact after normalisation)."""
import random

import numpy as np
import pytest

from oracle import orc
from tests import pyref
from tests.pyref import R_MOD
from tests.util import gpu_ctx, pkg, random_fr_mont, to_dev

pytestmark = pytest.mark.gpu


def affine(j):
    return pkg().api.jac_to_affine(j)


def make_bases(n, seed):
    return orc.fixed_base_batch(random_fr_mont(n, seed))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 31, 32, 33, 100, 1000])
def test_msm_small_vs_naive(n):
    ctx = gpu_ctx()
    rng = random.Random(n)
    bases = make_bases(max(n, 1), 10 + n)[:n]
    sc = [rng.randrange(R_MOD) for _ in range(n)]
    if n >= 4:
        sc[0], sc[1], sc[2] = 0, 1, R_MOD - 1
    if n >= 33:
        bases[5] = 0                      # identity base
        bases[7] = bases[6]               # duplicate base -> doubling inside a bucket when scalars match
        sc[7] = sc[6]
        bases[9] = bases[8]
        sc[9] = (R_MOD - sc[8]) % R_MOD   # P and -P in the same bucket -> identity
    S = orc.fr_from_ints(sc) if n else np.zeros((0, 4), dtype=np.uint64)
    got = affine(ctx.msm(S, bases))
    assert np.array_equal(got, orc.msm_naive(S, bases))


def test_msm_all_zero_and_identity_result():
    ctx = gpu_ctx()
    bases = make_bases(64, 3)
    S = np.zeros((64, 4), dtype=np.uint64)
    out = ctx.msm(S, bases)
    assert not affine(out).any()
    # halo2curves G1::identity() is (0, 1, 0)
    assert orc.fq_to_ints(out.reshape(3, 4)) == [0, 1, 0]


@pytest.mark.parametrize("log_n,kind", [(10, "uniform"), (12, "uniform"), (14, "bits"), (14, "bytes"), (14, "limbs64"), (16, "uniform"),
                                        (16, "mixed"), (17, "uniform")])
def test_msm_vs_best_multiexp(log_n, kind):
    ctx = gpu_ctx()
    n = 1 << log_n
    bases = make_bases(n, 40 + log_n)
    rng = np.random.default_rng(log_n)
    S = random_fr_mont(n, 50 + log_n)
    if kind != "uniform":
        small = np.zeros((n, 4), dtype=np.uint64)
        if kind == "bits":
            small[:, 0] = rng.integers(0, 2, n)
        elif kind == "bytes":
            small[:, 0] = rng.integers(0, 256, n)
        elif kind == "limbs64":
            small[:, 0] = rng.integers(0, 1 << 63, n, dtype=np.uint64)
        else:  # halo2-base like mixture: bits / bytes / limbs / uniform
            sel = rng.integers(0, 20, n)
            small[:, 0] = np.where(sel < 10, rng.integers(0, 2, n), np.where(sel < 15, rng.integers(0, 256, n),
                                                                            rng.integers(0, 1 << 63, n, dtype=np.uint64)))
        Sm = orc.field_op("fr", "from_canonical", small)
        if kind == "mixed":
            keep = rng.integers(0, 20, n) == 0
            Sm[keep] = S[keep]
        S = Sm
    got = affine(ctx.msm(S, bases))
    assert np.array_equal(got, orc.best_multiexp(S, bases))


def test_msm_batched_dev():
    ctx = gpu_ctx()
    n, ncols = 1 << 12, 5
    bases = make_bases(n, 77)
    S = random_fr_mont(n * ncols, 78)
    S[n:2 * n] = 0
    S[n + 5] = orc.fr_from_ints([1])[0]
    out = ctx.msm_dev(to_dev(S), to_dev(bases), n, ncols)
    for c in range(ncols):
        assert np.array_equal(affine(out[c:c + 1]), orc.best_multiexp(S[c * n:(c + 1) * n], bases)), c


@pytest.mark.parametrize("k", [4, 8, 11])
def test_srs_setup_matches_oracle(k):
    ctx = gpu_ctx()
    s = orc.fr_from_ints([pyref.ChaChaRng(bytes(32), 20).fr_random()])   # gen_srs secret (helpers.rs:210)
    params = pkg().ParamsKZG.setup(k, s, ctx=ctx)
    g, gl = orc.srs_setup(k, s)
    assert np.array_equal(params.get_g(0), g)
    assert np.array_equal(params.get_g(1), gl)
    n = 1 << k
    poly = random_fr_mont(n, 5 + k)
    assert np.array_equal(affine(params.commit(poly)), orc.best_multiexp(poly, g))
    assert np.array_equal(affine(params.commit_lagrange(poly)), orc.best_multiexp(poly, gl))
    # shorter polynomial (instance columns, h pieces): first `len` bases only
    assert np.array_equal(affine(params.commit(poly[: n // 2 + 1])), orc.best_multiexp(poly[: n // 2 + 1], g[: n // 2 + 1]))


def test_srs_load_and_batched_commit():
    ctx = gpu_ctx()
    k = 10
    n = 1 << k
    s = orc.fr_from_ints([123456789])
    g, gl = orc.srs_setup(k, s)
    params = pkg().ParamsKZG(k, g=g, g_lagrange=gl, ctx=ctx)
    polys = random_fr_mont(n * 4, 99)
    polys[2 * n:3 * n, 1:] = 0
    polys[2 * n:3 * n, 0] &= np.uint64(1)      # raw small Montgomery words: still valid field elements
    for basis, bases in ((0, g), (1, gl)):
        out = params.commit_dev(to_dev(polys), n, 4, basis)
        for c in range(4):
            assert np.array_equal(affine(out[c:c + 1]), orc.best_multiexp(polys[c * n:(c + 1) * n], bases)), (basis, c)


def test_commit_coeff_equals_commit_lagrange_k17():
    """Size-independent property at BASELINE config-1 size: commit(p) in the coefficient basis equals
    commit_lagrange of its evaluations, and equals [p(s)]G."""
    ctx = gpu_ctx()
    k = 17
    n = 1 << k
    s_int = pyref.ChaChaRng(bytes(32), 20).fr_random()
    params = pkg().ParamsKZG.setup(k, orc.fr_from_ints([s_int]), ctx=ctx)
    dom = pkg().EvaluationDomain(4, k, ctx=ctx)
    coeffs = random_fr_mont(n, 4242)
    lag = dom.coeff_to_lagrange(coeffs)
    c1 = affine(params.commit(coeffs))
    c2 = affine(params.commit_lagrange(lag))
    assert np.array_equal(c1, c2)
    # [p(s)]G with p(s) evaluated by Python big ints on a sparse polynomial for speed
    sparse = np.zeros((n, 4), dtype=np.uint64)
    idx = [0, 1, 77, n - 1]
    vals = [5, R_MOD - 3, 1 << 200, 99]
    sparse[idx] = orc.fr_from_ints(vals)
    ps = sum(v * pow(s_int, i, R_MOD) for i, v in zip(idx, vals)) % R_MOD
    assert orc.g1_to_ints(affine(params.commit(sparse)))[0] == pyref.ec_mul(pyref.G1_GEN, ps)


def test_published_eip196_known_answer():
    """EIP-196 (alt_bn128) ecAdd vector (1, 2) + (1, 2), through zkc_msm_g1 with scalars [1, 1] and [2] on the generator"""
    from tests.test_oracle_ops import EIP196_2G
    p = pkg()
    ctx = gpu_ctx()
    G = orc.g1_generator()
    both = np.concatenate([G, G])
    assert orc.g1_to_ints(affine(ctx.msm(orc.fr_from_ints([1, 1]), both)))[0] == EIP196_2G
    assert orc.g1_to_ints(affine(ctx.msm(orc.fr_from_ints([2]), G)))[0] == EIP196_2G


def test_params_file_write_read_on_device(tmp_path):
    """ParamsKZG::write / read through kzg_bn254_{k}.srs (RawBytes): device SRS -> file -> device SRS, same commitments"""
    from tests.circuits import SRS_SECRET
    p = pkg()
    ctx = gpu_ctx()
    k = 8
    params = p.ParamsKZG.setup(k, orc.fr_from_ints([SRS_SECRET]), ctx=ctx)
    s_g2 = p.api.g2_mul(p.api.g2_generator(), orc.fr_from_ints([SRS_SECRET]))
    path = str(tmp_path / ("kzg_bn254_%d.srs" % k))
    assert params.write(path, s_g2) == 4 + 2 * 64 * (1 << k) + 256
    back, g2, s_g2_back = p.ParamsKZG.read(path, ctx=ctx)
    assert back.k == k and np.array_equal(back.get_g(0), params.get_g(0)) and np.array_equal(back.get_g(1), params.get_g(1))
    assert np.array_equal(g2, p.api.g2_generator()) and np.array_equal(s_g2_back, s_g2)
    poly = random_fr_mont(1 << k, 3)
    assert np.array_equal(back.commit(poly), params.commit(poly)) and np.array_equal(back.commit_lagrange(poly), params.commit_lagrange(poly))
    # e(g[1], G2) == e(g[0], s_g2): the file's G1 and G2 halves belong to the same secret
    g = back.get_g(0)
    neg_g0 = orc.g1_from_ints([pyref.ec_neg(orc.g1_to_ints(g[:1])[0])])
    assert p.api.pairing_check(np.concatenate([g[1:2], neg_g0]), np.concatenate([g2, s_g2_back]))
