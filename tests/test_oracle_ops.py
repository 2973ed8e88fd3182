"""Pins the CPU oracle (oracle/) against Python big-integer ground truth and the known answers in
SURVEY.md §8(c)-4.  The reference's own tests hold no golden vector at this boundary,
so these uniqueness checks are what the oracle stands on: parity unpinned."""
import random

import numpy as np
import pytest

from oracle import orc
from tests import pyref
from tests.pyref import R_MOD, P_MOD


def rnd_fr(rng, n):
    return [rng.randrange(R_MOD) for _ in range(n)]


def test_field_constants():
    assert pow(pyref.ROOT_OF_UNITY, 1 << 28, R_MOD) == 1 and pow(pyref.ROOT_OF_UNITY, 1 << 27, R_MOD) != 1
    assert pyref.ROOT_OF_UNITY == pow(7, (R_MOD - 1) >> 28, R_MOD)
    assert pyref.DELTA == pow(7, 1 << 28, R_MOD)
    assert pow(pyref.ZETA, 3, R_MOD) == 1 and pyref.ZETA != 1
    # SURVEY §8a a5 omega prefixes/suffixes
    for k, hi, lo in [(17, 0x304cd1e7, 0x79c2c3ea), (19, 0x0cf1526a, 0x568fc082), (22, 0x18c95f1a, 0x8326bede)]:
        w = pyref.omega_for(k)
        assert w >> 224 == hi and w & 0xFFFFFFFF == lo


@pytest.mark.parametrize("which,mod", [("fr", R_MOD), ("fq", P_MOD)])
def test_field_ops(which, mod):
    rng = random.Random(1)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 64) - 1, 1 << 64]
    a = edge + [rng.randrange(mod) for _ in range(200)]
    b = list(reversed(edge)) + [rng.randrange(mod) for _ in range(200)]
    conv = orc.fr_from_ints if which == "fr" else orc.fq_from_ints
    back = orc.fr_to_ints if which == "fr" else orc.fq_to_ints
    A, B = conv(a), conv(b)
    assert back(orc.field_op(which, "add", A, B)) == [(x + y) % mod for x, y in zip(a, b)]
    assert back(orc.field_op(which, "sub", A, B)) == [(x - y) % mod for x, y in zip(a, b)]
    assert back(orc.field_op(which, "mul", A, B)) == [(x * y) % mod for x, y in zip(a, b)]
    assert back(orc.field_op(which, "neg", A)) == [(-x) % mod for x in a]
    assert back(orc.field_op(which, "inv", A)) == [pow(x, -1, mod) if x else 0 for x in a]
    # Montgomery conversion done by the C code agrees with the Python one
    assert np.array_equal(orc.field_op(which, "from_canonical", orc.ints_to_limbs(a)), A)
    assert orc.limbs_to_ints(orc.field_op(which, "to_canonical", A)) == a


def test_from_u512():
    rng = random.Random(2)
    vals = [0, 1, (1 << 512) - 1, R_MOD << 256] + [rng.getrandbits(512) for _ in range(50)]
    got = orc.fr_to_ints(orc.field_op("fr", "from_u512", orc.ints_to_limbs(vals, 8)))
    assert got == [v % R_MOD for v in vals]


def test_chacha_known_answers():
    # ChaCha20 zero-key keystream and the gen_srs secret s
    ks = pyref.ChaChaRng(bytes(32), 20)
    stream = b"".join(ks.next_u32().to_bytes(4, "little") for _ in range(16))
    assert stream.hex() == ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                            "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    s = pyref.ChaChaRng(bytes(32), 20).fr_random()
    assert s == 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba
    assert pyref.seed_from_u64(0).hex() == "ecf273f981b5cd4587f0467306ad6cadd0d0a3e33317e767f29bea72d78a7dfe"
    # independent ChaCha20: `cryptography` (IETF layout: 32-bit counter + 96-bit nonce, same block 0..)
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    enc = Cipher(algorithms.ChaCha20(bytes(32), bytes(16)), mode=None).encryptor()
    assert enc.update(bytes(64)) == stream


EIP196_2G = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
             9918110051302171585080402603319702774565515993150576347155970296011118125764)


def test_g1_ops():
    rng = random.Random(3)
    G = orc.g1_generator()
    assert orc.g1_to_ints(G) == [pyref.G1_GEN]
    for s in [0, 1, 2, 3, R_MOD - 1, rng.randrange(R_MOD), rng.randrange(R_MOD)]:
        got = orc.g1_to_ints(orc.g1_mul(G, orc.fr_from_ints([s])))[0]
        assert got == pyref.ec_mul(pyref.G1_GEN, s)
    P = pyref.ec_mul(pyref.G1_GEN, 12345)
    Q = pyref.ec_mul(pyref.G1_GEN, 99999)
    for a, b in [(P, Q), (P, P), (P, pyref.ec_neg(P)), (P, None), (None, Q), (None, None)]:
        got = orc.g1_to_ints(orc.g1_add(orc.g1_from_ints([a]), orc.g1_from_ints([b])))[0]
        assert got == pyref.ec_add(a, b)
    assert orc.g1_on_curve(orc.g1_from_ints([P]))
    assert not orc.g1_on_curve(orc.g1_from_ints([(P[0], P[1] + 1)]))
    # published known answer: EIP-196 (alt_bn128 ecAdd) (1, 2) + (1, 2)
    assert orc.g1_to_ints(orc.g1_mul(G, orc.fr_from_ints([2])))[0] == EIP196_2G
    assert orc.g1_to_ints(orc.g1_add(G, G))[0] == EIP196_2G


def _points(n, seed):
    rng = random.Random(seed)
    ks = [rng.randrange(1, R_MOD) for _ in range(n)]
    arr = orc.fixed_base_batch(orc.fr_from_ints(ks))
    return ks, arr


def test_fixed_base_and_msm_small():
    ks, bases = _points(40, 4)
    pts = orc.g1_to_ints(bases)
    assert pts[:5] == [pyref.ec_mul(pyref.G1_GEN, k) for k in ks[:5]]
    rng = random.Random(5)
    for n in [0, 1, 3, 4, 31, 32, 40]:
        sc = [rng.randrange(R_MOD) for _ in range(n)]
        if n >= 4:
            sc[0] = 0; sc[1] = 1; sc[2] = R_MOD - 1
        want = pyref.ec_msm(sc, pts[:n])
        S = orc.fr_from_ints(sc) if n else np.zeros((0, 4), dtype=np.uint64)
        assert orc.g1_to_ints(orc.msm_naive(S, bases[:n]))[0] == want
        for threads in (1, 3, 8):
            assert orc.g1_to_ints(orc.best_multiexp(S, bases[:n], threads))[0] == want


def test_best_multiexp_matches_naive_2k():
    ks, bases = _points(2048, 6)
    rng = random.Random(7)
    sc = [rng.randrange(R_MOD) for _ in range(2048)]
    for i in range(0, 2048, 7):
        sc[i] = rng.randrange(2)          # bit-valued witness cells
    for i in range(3, 2048, 11):
        sc[i] = rng.randrange(1 << 64)    # limb-valued cells
    bases[5] = 0                           # an identity base
    S = orc.fr_from_ints(sc)
    want = orc.msm_naive(S, bases)
    assert np.array_equal(orc.best_multiexp(S, bases, 1), want)
    assert np.array_equal(orc.best_multiexp(S, bases, 8), want)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8])
def test_best_fft_vs_naive_dft(log_n):
    rng = random.Random(8 + log_n)
    n = 1 << log_n
    a = rnd_fr(rng, n)
    w = pyref.omega_for(log_n)
    want = pyref.dft(a, w)
    for threads in (1, 4):
        got = orc.fr_to_ints(orc.best_fft(orc.fr_from_ints(a), orc.fr_from_ints([w]), log_n, threads))
        assert got == want


def test_fft_roundtrip_and_domain():
    rng = random.Random(20)
    j, k = 4, 9
    n = 1 << k
    d = orc.domain_constants(j, k)
    assert d["extended_k"] == k + 2
    assert orc.fr_to_ints(d["omega"])[0] == pyref.omega_for(k)
    assert orc.fr_to_ints(d["extended_omega"])[0] == pyref.omega_for(k + 2)
    assert orc.fr_to_ints(d["g_coset"])[0] == pyref.ZETA
    assert orc.fr_to_ints(d["g_coset_inv"])[0] == pow(pyref.ZETA, 2, R_MOD)
    coeffs = rnd_fr(rng, n)
    A = orc.fr_from_ints(coeffs)
    lag = orc.coeff_to_lagrange(j, k, A)
    w = pyref.omega_for(k)
    lag_i = orc.fr_to_ints(lag)
    for i in (0, 1, 5, n - 1):
        assert lag_i[i] == pyref.poly_eval(coeffs, pow(w, i, R_MOD))
    assert np.array_equal(orc.lagrange_to_coeff(j, k, lag), A)
    # coeff_to_extended evaluates on zeta * extended_omega^i
    ext = orc.coeff_to_extended(j, k, A)
    we = pyref.omega_for(k + 2)
    ext_i = orc.fr_to_ints(ext)
    for i in (0, 1, 2, 3, 777, 4 * n - 1):
        assert ext_i[i] == pyref.poly_eval(coeffs, pyref.ZETA * pow(we, i, R_MOD) % R_MOD)
    back = orc.extended_to_coeff(j, k, ext)
    assert np.array_equal(back[:n], A) and not back[n:].any()
    # divide_by_vanishing_poly multiplies row i by 1 / ((zeta*we^i)^n - 1)
    dv = orc.fr_to_ints(orc.divide_by_vanishing(j, k, ext))
    for i in (0, 1, 2, 3, 4, 1001):
        x = pyref.ZETA * pow(we, i, R_MOD) % R_MOD
        assert dv[i] == ext_i[i] * pow(pow(x, n, R_MOD) - 1, -1, R_MOD) % R_MOD


def test_srs_setup_small():
    k = 5
    n = 1 << k
    s = pyref.ChaChaRng(bytes(32), 20).fr_random()   # the gen_srs secret
    g, gl = orc.srs_setup(k, orc.fr_from_ints([s]))
    gi = orc.g1_to_ints(g)
    assert gi[0] == pyref.G1_GEN and gi[3] == pyref.ec_mul(pyref.G1_GEN, pow(s, 3, R_MOD))
    # commit(coeff) == commit_lagrange(lagrange) for the same polynomial
    rng = random.Random(9)
    coeffs = rnd_fr(rng, n)
    A = orc.fr_from_ints(coeffs)
    lag = orc.coeff_to_lagrange(2, k, A)
    c1 = orc.best_multiexp(A, g, 2)
    c2 = orc.best_multiexp(lag, gl, 2)
    assert np.array_equal(c1, c2)
    assert orc.g1_to_ints(c1)[0] == pyref.ec_mul(pyref.G1_GEN, pyref.poly_eval(coeffs, s))


def test_oracle_chacha_stream_matches_python_restatement():
    seed = pyref.seed_from_u64(77)
    slow = pyref.ChaChaRng(seed, 20)
    want = [slow.fr_random() for _ in range(40)]
    fast = orc.ChaCha20Rng(seed)
    assert [fast.fr_random() for _ in range(3)] == want[:3]
    assert orc.fr_to_ints(fast.fr_random_bulk(37)) == want[3:]
    assert orc.ChaCha20Rng(bytes(32)).fr_random() == 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba


def test_std_rng_chacha12_stream():
    seed = pyref.seed_from_u64(5)
    slow = pyref.ChaChaRng(seed, 12)
    want = [slow.fr_random() for _ in range(6)]
    assert orc.fr_to_ints(orc.ChaCha20Rng(seed, rounds=12).fr_random_bulk(6)) == want
    assert want != [pyref.ChaChaRng(seed, 20).fr_random() for _ in range(6)]


def test_residue_class_identities_of_the_extended_coset():
    """The algebra behind the class-major prover (csrc/ntt.cu dom_coeff_to_classes / dom_classes_to_pieces), checked on the
    oracle's own transforms with Python integers in between — no GPU involved:
      (1) rows c, c + 2^e, c + 2*2^e, ... of coeff_to_extended are ONE size-n transform of the coefficients scaled by
          zeta^(i mod 3) * w_ext^(c i)  (the four-step split N1 = n, N2 = 2^e of a zero-padded input);
      (2) a polynomial with fewer than q n coefficients is determined by its values on q classes: per-class coset iNTT, then the
          inverse of the q x q Vandermonde matrix in tau_c = (zeta w_ext^c)^n gives extended_to_coeff's output piece by piece."""
    import random
    R = pyref.R_MOD
    rnd = random.Random(11)
    for j, k in [(4, 5), (5, 4), (3, 6)]:
        d = orc.domain_constants(j, k)
        ek = d["extended_k"]
        n, en, e = 1 << k, 1 << ek, ek - k
        w_ext, w, zeta = orc.fr_to_ints(d["extended_omega"])[0], orc.fr_to_ints(d["omega"])[0], orc.fr_to_ints(d["g_coset"])[0]
        assert pow(w_ext, 1 << e, R) == w and pow(zeta, 3, R) == 1
        # (1) forward
        a = [rnd.randrange(R) for _ in range(n)]
        ext = orc.fr_to_ints(orc.coeff_to_extended(j, k, orc.fr_from_ints(a)))
        for c in range(1 << e):
            scaled = [a[i] * pow(zeta, i % 3, R) * pow(w_ext, c * i, R) % R for i in range(n)]
            cls = orc.fr_to_ints(orc.best_fft(orc.fr_from_ints(scaled), d["omega"], k))
            assert cls == ext[c::1 << e], (j, k, c)
        # (2) the way back from q = j - 1 classes
        q = j - 1
        h = [rnd.randrange(R) for _ in range(q * n)]
        hs = [h[i] * pow(zeta, i % 3, R) % R for i in range(q * n)] + [0] * (en - q * n)
        vals = orc.fr_to_ints(orc.best_fft(orc.fr_from_ints(hs), d["extended_omega"], ek))       # h on the whole extended coset
        assert orc.fr_to_ints(orc.extended_to_coeff(j, k, orc.fr_from_ints(vals)))[:q * n] == h
        ninv, g, tau = pow(n, -1, R), [], []
        for c in range(q):
            zc = zeta * pow(w_ext, c, R) % R
            y = orc.fr_to_ints(orc.best_fft(orc.fr_from_ints(vals[c::1 << e]), d["omega_inv"], k))
            g.append([y[i] * ninv * pow(zc, -i, R) % R for i in range(n)])
            tau.append(pow(zc, n, R))
        # solve V H = g with V[c][p] = tau_c^p (Gauss-Jordan over the field, per coefficient index all at once)
        V = [[pow(tau[c], p, R) for p in range(q)] + [1 if p == c else 0 for p in range(q)] for c in range(q)]
        for col in range(q):
            piv = next(r for r in range(col, q) if V[r][col])
            V[col], V[piv] = V[piv], V[col]
            inv = pow(V[col][col], -1, R)
            V[col] = [x * inv % R for x in V[col]]
            for r in range(q):
                if r != col and V[r][col]:
                    f = V[r][col]
                    V[r] = [(x - f * y) % R for x, y in zip(V[r], V[col])]
        Minv = [row[q:] for row in V]
        for p in range(q):
            piece = [sum(Minv[p][c] * g[c][i] for c in range(q)) % R for i in range(n)]
            assert piece == h[p * n:(p + 1) * n], (j, k, p)
