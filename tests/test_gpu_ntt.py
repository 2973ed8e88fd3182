"""GPU parity: best_fft / EvaluationDomain through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from oracle import orc
from tests import pyref
from tests.util import gpu_ctx, pkg, random_fr_mont, to_dev, to_host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 10, 11, 12, 13, 15, 16, 17, 18, 19, 20, 21])
def test_best_fft_matches_oracle(log_n):
    ctx = gpu_ctx()
    a = random_fr_mont(1 << log_n, 100 + log_n)
    w = orc.fr_from_ints([pyref.omega_for(log_n)])
    winv = orc.fr_from_ints([pow(pyref.omega_for(log_n), -1, pyref.R_MOD)])
    got = ctx.fft(a, w, log_n)
    assert np.array_equal(got, orc.best_fft(a, w, log_n))
    got_inv = ctx.fft(a, winv, log_n)
    assert np.array_equal(got_inv, orc.best_fft(a, winv, log_n))


def test_best_fft_rejects_bad_omega():
    ctx = gpu_ctx()
    a = random_fr_mont(16, 1)
    with pytest.raises(pkg().ZkcError):
        ctx.fft(a, orc.fr_from_ints([5]), 4)


def test_best_fft_edge_values():
    ctx = gpu_ctx()
    log_n = 12
    n = 1 << log_n
    w = orc.fr_from_ints([pyref.omega_for(log_n)])
    zeros = np.zeros((n, 4), dtype=np.uint64)
    assert not ctx.fft(zeros, w, log_n).any()
    ones = np.repeat(orc.fr_from_ints([1]), n, axis=0)
    got = orc.fr_to_ints(ctx.fft(ones, w, log_n))
    assert got[0] == n and not any(got[1:])
    maxv = np.repeat(orc.fr_from_ints([pyref.R_MOD - 1]), n, axis=0)
    assert np.array_equal(ctx.fft(maxv, w, log_n), orc.best_fft(maxv, w, log_n))


@pytest.mark.parametrize("log_n,ncols", [(9, 5), (13, 3), (17, 4), (19, 2)])
def test_batched_fft_dev(log_n, ncols):
    import torch
    ctx = gpu_ctx()
    n = 1 << log_n
    a = random_fr_mont(n * ncols, 7 + log_n)
    w = orc.fr_from_ints([pyref.omega_for(log_n)])
    t = to_dev(a)
    ctx.fft_dev(t, w, log_n, ncols)
    ctx.sync()
    got = to_host(t)
    for c in range(ncols):
        assert np.array_equal(got[c * n:(c + 1) * n], orc.best_fft(a[c * n:(c + 1) * n], w, log_n))


@pytest.mark.parametrize("j,k", [(4, 4), (4, 9), (4, 12), (3, 13), (4, 15), (4, 17), (5, 14), (9, 10)])
def test_domain_conversions(j, k):
    ctx = gpu_ctx()
    dom = pkg().EvaluationDomain(j, k, ctx=ctx)
    d = orc.domain_constants(j, k)
    assert dom.extended_k == d["extended_k"]
    for name in orc.DOMAIN_FIELDS:
        assert np.array_equal(getattr(dom, name), d[name]), name
    a = random_fr_mont(1 << k, 33 + k)
    coeff = dom.lagrange_to_coeff(a)
    assert np.array_equal(coeff, orc.lagrange_to_coeff(j, k, a))
    assert np.array_equal(dom.coeff_to_lagrange(coeff), a)
    ext = dom.coeff_to_extended(coeff)
    assert np.array_equal(ext, orc.coeff_to_extended(j, k, coeff))
    # extended_to_coeff on an arbitrary extended vector (not only on images of coeff_to_extended)
    e = random_fr_mont(1 << dom.extended_k, 55 + k)
    assert np.array_equal(dom.extended_to_coeff(e), orc.extended_to_coeff(j, k, e))
    back = dom.extended_to_coeff(ext)
    assert np.array_equal(back[: 1 << k], coeff) and not back[1 << k:].any()
    # divide_by_vanishing_poly (device-resident API)
    t = to_dev(e)
    dom.divide_by_vanishing_poly_dev(t)
    ctx.sync()
    assert np.array_equal(to_host(t), orc.divide_by_vanishing(j, k, e))


@pytest.mark.parametrize("j,k,ncols", [(4, 4, 2), (4, 9, 3), (3, 13, 2), (4, 15, 2), (4, 17, 3), (5, 14, 1), (9, 10, 2), (2, 8, 2)])
def test_coeff_to_extended_by_residue_class(j, k, ncols):
    """The class split of the extended coset (ntt.cu dom_coeff_to_classes; what team proving shards by): class c of a column is
    one size-n transform and equals rows c, c + 2^e, c + 2 * 2^e, ... of the oracle's coeff_to_extended; classes outside
    [c0, c1) are left untouched; the natural-order permutation gives the oracle's column back."""
    import torch
    ctx = gpu_ctx()
    dom = pkg().EvaluationDomain(j, k, ctx=ctx)
    n, en = 1 << k, 1 << dom.extended_k
    ncls = en // n
    a = random_fr_mont(n * ncols, 700 + k)
    want = [orc.coeff_to_extended(j, k, a[c * n:(c + 1) * n]) for c in range(ncols)]
    src = to_dev(a)
    for c0, c1 in sorted({(0, ncls), (0, 1), (ncls - 1, ncls), (ncls // 2, ncls)}):
        dst = torch.full((en * ncols, 4), -1, dtype=torch.int64, device="cuda")
        dom.coeff_to_extended_classes_dev(src, dst, ncols, c0, c1)
        ctx.sync()
        got = to_host(dst)
        for col in range(ncols):
            for c in range(ncls):
                blk = got[col * en + c * n: col * en + (c + 1) * n]
                if c0 <= c < c1:
                    assert np.array_equal(blk, want[col][c::ncls]), (col, c)
                else:
                    assert (blk == np.uint64(0xFFFFFFFFFFFFFFFF)).all(), (col, c)
    nat = torch.empty((en, 4), dtype=torch.int64, device="cuda")
    dom.extended_classes_to_natural_dev(dst[en * (ncols - 1):], nat)   # last loop iteration transformed a suffix of the classes only
    dst = torch.empty((en * ncols, 4), dtype=torch.int64, device="cuda")
    dom.coeff_to_extended_classes_dev(src, dst, ncols, 0, ncls)
    dom.extended_classes_to_natural_dev(dst[en * (ncols - 1):], nat)
    ctx.sync()
    assert np.array_equal(to_host(nat), want[ncols - 1])


def test_domain_batched_dev_full_size():
    """BASELINE config-1 shape: k=17, extended 2^19, a few columns; round trip + spot parity."""
    ctx = gpu_ctx()
    j, k, ncols = 4, 17, 3
    dom = pkg().EvaluationDomain(j, k, ctx=ctx)
    n, en = 1 << k, 1 << dom.extended_k
    a = random_fr_mont(n * ncols, 2024)
    import torch
    src = to_dev(a)
    dst = torch.empty((en * ncols, 4), dtype=torch.int64, device="cuda")
    dom.coeff_to_extended_dev(src, dst, ncols)
    ctx.sync()
    ext = to_host(dst)
    assert np.array_equal(ext[en:2 * en], orc.coeff_to_extended(j, k, a[n:2 * n]))
    dom.extended_to_coeff_dev(dst, ncols)
    ctx.sync()
    back = to_host(dst)
    for c in range(ncols):
        assert np.array_equal(back[c * en:c * en + n], a[c * n:(c + 1) * n])
        assert not back[c * en + n:(c + 1) * en].any()


def test_field_vec_ops():
    import torch
    ctx = gpu_ctx()
    n = 5000
    for field in ("fr", "fq"):
        a, b = random_fr_mont(n, 1), random_fr_mont(n, 2)
        a[0] = 0
        a[17] = 0
        ta, tb = to_dev(a), to_dev(b)
        out = torch.empty_like(ta)
        for op in ("add", "sub", "mul"):
            ctx.field_vec_op_dev(field, op, ta, tb, out)
            ctx.sync()
            assert np.array_equal(to_host(out), orc.field_op(field, op, a, b)), (field, op)
        for op in ("inv", "neg", "from_canonical", "to_canonical"):
            ctx.field_vec_op_dev(field, op, ta, None, out)
            ctx.sync()
            assert np.array_equal(to_host(out), orc.field_op(field, op, a)), (field, op)


def test_field_square_edge_values():
    """dedicated squaring (ff.cuh fe_sqr: 36 + 64 wide multiplies) == a * a, on random and carry-heavy inputs, both fields"""
    import torch
    ctx = gpu_ctx()
    P = {"fr": pyref.R_MOD, "fq": pyref.P_MOD}
    for field in ("fr", "fq"):
        p = P[field]
        edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, (1 << 253) - 1, (1 << 253), ((1 << 254) - 1) % p,
                0xFFFFFFFF, (1 << 32), (1 << 224) - 1, int("f" * 56, 16) % p, int("ffffffff00000000" * 4, 16) % p,
                int("00000000ffffffff" * 4, 16) % p, p - (1 << 32), p - (1 << 128)]
        # the kernel sees MONTGOMERY words: feed raw limb patterns directly too (every limb all-ones below p is impossible, so
        # use the largest raw values: p - 1, p - 2^32k, and words with all-ones low limbs)
        raw = np.zeros((len(edge), 4), dtype=np.uint64)
        for i, v in enumerate(edge):
            for l in range(4):
                raw[i, l] = (v >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
        a = np.concatenate([raw, random_fr_mont(20000, 9) if field == "fr" else orc.field_op("fq", "from_canonical", random_fr_mont(20000, 9))])
        ta = to_dev(a)
        out = torch.empty_like(ta)
        ctx.field_vec_op_dev(field, "square", ta, None, out)
        ctx.sync()
        assert np.array_equal(to_host(out), orc.field_op(field, "mul", a, a)), field


def test_batch_invert_assigned():
    import torch
    ctx = gpu_ctx()
    n = 3000
    num, den = random_fr_mont(n, 5), random_fr_mont(n, 6)
    den[3] = 0
    den[10] = orc.fr_from_ints([1])[0]
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    ctx.batch_invert_assigned_dev(to_dev(num), to_dev(den), out)
    ctx.sync()
    want = orc.field_op("fr", "mul", num, orc.field_op("fr", "inv", den))
    assert np.array_equal(to_host(out), want)
    assert not to_host(out)[3].any() and np.array_equal(to_host(out)[10], num[10])


def test_large_batch_inversion_and_scan_paths():
    """n >= 2^15 takes the two-scan batch inversion; n > 16 * 32768 takes the recursive scan"""
    import torch
    ctx = gpu_ctx()
    n = (1 << 20) + 12345
    a = random_fr_mont(n, 77)
    a[0] = 0
    a[n - 1] = 0
    a[4097] = 0
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    ctx.field_vec_op_dev("fr", "inv", to_dev(a), None, out)
    ctx.sync()
    got = to_host(out)
    assert np.array_equal(got, orc.batch_invert(a))
