"""ctypes binding of the CPU oracle (oracle/_build/libzkc_oracle.so).  TEST INFRASTRUCTURE ONLY.

Field elements travel as numpy uint64 arrays of shape (n, 4): little-endian limbs of the Montgomery
residue, the same bytes halo2curves' `Fr`/`Fq` hold.  G1Affine = (n, 8).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzkc_oracle.so")

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
MONT_R = 1 << 256


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_domain_constants.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- conversions between Python ints and limb arrays -----------------------------------------
def ints_to_limbs(vals, words=4):
    out = np.zeros((len(vals), words), dtype=np.uint64)
    for i, v in enumerate(vals):
        for w in range(words):
            out[i, w] = (v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, arr.shape[-1])
    return [sum(int(arr[i, w]) << (64 * w) for w in range(arr.shape[1])) for i in range(arr.shape[0])]


def fr_from_ints(vals):
    """canonical ints -> Montgomery limb array (done in Python, independent of the C code)."""
    return ints_to_limbs([(v % R_MOD) * MONT_R % R_MOD for v in vals])


def fr_to_ints(arr):
    rinv = pow(MONT_R, -1, R_MOD)
    return [v * rinv % R_MOD for v in limbs_to_ints(arr)]


def fq_from_ints(vals):
    return ints_to_limbs([(v % P_MOD) * MONT_R % P_MOD for v in vals])


def fq_to_ints(arr):
    rinv = pow(MONT_R, -1, P_MOD)
    return [v * rinv % P_MOD for v in limbs_to_ints(arr)]


def g1_from_ints(pts):
    """[(x, y) | None] -> (n, 8) Montgomery array; None = identity = (0, 0)."""
    flat = []
    for pt in pts:
        x, y = (0, 0) if pt is None else pt
        flat += [x, y]
    return fq_from_ints(flat).reshape(-1, 8)


def g1_to_ints(arr):
    v = fq_to_ints(np.ascontiguousarray(arr).reshape(-1, 4))
    out = []
    for i in range(0, len(v), 2):
        out.append(None if (v[i] == 0 and v[i + 1] == 0) else (v[i], v[i + 1]))
    return out


# ---- operators ---------------------------------------------------------------------------------
OPS = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "from_canonical": 4, "to_canonical": 5, "neg": 6, "from_u512": 7}


def field_op(which, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    n = a.shape[0]
    out = np.empty((n, 4), dtype=np.uint64)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    lib().orc_field_op(C.c_int(0 if which == "fr" else 1), C.c_int(OPS[op]), _p(a), None if bb is None else _p(bb), _p(out), C.c_size_t(n))
    return out


def g1_generator():
    out = np.empty((1, 8), dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_mul(p, s):
    out = np.empty((1, 8), dtype=np.uint64)
    lib().orc_g1_mul(_p(np.ascontiguousarray(p)), _p(np.ascontiguousarray(s)), _p(out))
    return out


def g1_add(p, q):
    out = np.empty((1, 8), dtype=np.uint64)
    lib().orc_g1_add(_p(np.ascontiguousarray(p)), _p(np.ascontiguousarray(q)), _p(out))
    return out


def g1_on_curve(p):
    return bool(lib().orc_g1_on_curve(_p(np.ascontiguousarray(p))))


def g1_to_affine(jac):
    jac = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 12)
    out = np.empty((jac.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_to_affine(_p(jac), _p(out), C.c_size_t(jac.shape[0]))
    return out


def msm_naive(scalars, bases):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.empty((1, 8), dtype=np.uint64)
    lib().orc_msm_naive(_p(scalars), _p(bases), C.c_size_t(scalars.shape[0]), _p(out))
    return out


def best_multiexp(scalars, bases, threads=0):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.empty((1, 8), dtype=np.uint64)
    lib().orc_best_multiexp(_p(scalars), _p(bases), C.c_size_t(scalars.shape[0]), C.c_int(threads), _p(out))
    return out


def best_fft(a, omega, log_n, threads=0):
    """in place on a copy; returns the transformed array"""
    a = np.array(a, dtype=np.uint64, copy=True)
    lib().orc_best_fft(_p(a), _p(np.ascontiguousarray(omega, dtype=np.uint64)), C.c_uint32(log_n), C.c_int(threads))
    return a


DOMAIN_FIELDS = ["omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv", "ifft_divisor", "extended_ifft_divisor"]


def domain_constants(j, k, zeta_choice=0):
    out = np.empty((8, 4), dtype=np.uint64)
    ek = lib().orc_domain_constants(C.c_uint32(j), C.c_uint32(k), C.c_int(zeta_choice), _p(out))
    d = {name: out[i:i + 1].copy() for i, name in enumerate(DOMAIN_FIELDS)}
    d["extended_k"] = int(ek)
    d["k"] = k
    d["j"] = j
    return d


def lagrange_to_coeff(j, k, a, threads=0):
    a = np.array(a, dtype=np.uint64, copy=True)
    lib().orc_lagrange_to_coeff(C.c_uint32(j), C.c_uint32(k), _p(a), C.c_int(threads))
    return a


def coeff_to_lagrange(j, k, a, threads=0):
    a = np.array(a, dtype=np.uint64, copy=True)
    lib().orc_coeff_to_lagrange(C.c_uint32(j), C.c_uint32(k), _p(a), C.c_int(threads))
    return a


def coeff_to_extended(j, k, a, zeta_choice=0, threads=0):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    ek = domain_constants(j, k)["extended_k"]
    out = np.empty((1 << ek, 4), dtype=np.uint64)
    lib().orc_coeff_to_extended(C.c_uint32(j), C.c_uint32(k), C.c_int(zeta_choice), _p(a), _p(out), C.c_int(threads))
    return out


def extended_to_coeff(j, k, a, zeta_choice=0, threads=0):
    a = np.array(a, dtype=np.uint64, copy=True)
    lib().orc_extended_to_coeff(C.c_uint32(j), C.c_uint32(k), C.c_int(zeta_choice), _p(a), C.c_int(threads))
    return a


def divide_by_vanishing(j, k, a, zeta_choice=0, threads=0):
    a = np.array(a, dtype=np.uint64, copy=True)
    lib().orc_divide_by_vanishing(C.c_uint32(j), C.c_uint32(k), C.c_int(zeta_choice), _p(a), C.c_int(threads))
    return a


def fixed_base_batch(scalars):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    out = np.empty((scalars.shape[0], 8), dtype=np.uint64)
    lib().orc_fixed_base_batch(_p(scalars), C.c_size_t(scalars.shape[0]), _p(out))
    return out


def srs_setup(k, s_mont, want_lagrange=True):
    n = 1 << k
    g = np.empty((n, 8), dtype=np.uint64)
    gl = np.empty((n, 8), dtype=np.uint64) if want_lagrange else None
    lib().orc_srs_setup(C.c_uint32(k), _p(np.ascontiguousarray(s_mont, dtype=np.uint64)), _p(g), None if gl is None else _p(gl))
    return g, gl


# ---- polynomial helpers (orc_poly.cpp) ---------------------------------------------------------
def vec_scalar(op, a, s):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fr_vec_scalar(C.c_int({"add": 0, "sub": 1, "mul": 2, "rsub": 3}[op]), _p(a), _p(np.ascontiguousarray(s, dtype=np.uint64)), _p(out),
                            C.c_size_t(a.shape[0]))
    return out


def vec_vec(op, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    assert a.shape == b.shape
    out = np.empty_like(a)
    lib().orc_fr_vec_vec(C.c_int({"add": 0, "sub": 1, "mul": 2}[op]), _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def batch_invert(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fr_batch_invert(_p(a), _p(out), C.c_size_t(a.shape[0]))
    return out


def eval_poly(coeffs, x):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    out = np.empty((1, 4), dtype=np.uint64)
    lib().orc_eval_poly(_p(coeffs), C.c_size_t(coeffs.shape[0]), _p(np.ascontiguousarray(x, dtype=np.uint64)), _p(out))
    return out


def kate_division(a, z):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.zeros((max(a.shape[0] - 1, 0), 4), dtype=np.uint64)
    lib().orc_kate_division(_p(a), C.c_size_t(a.shape[0]), _p(np.ascontiguousarray(z, dtype=np.uint64)), _p(out))
    return out


def prefix_product(r, start):
    r = np.ascontiguousarray(r, dtype=np.uint64)
    out = np.empty_like(r)
    lib().orc_prefix_product(_p(r), C.c_size_t(r.shape[0]), _p(np.ascontiguousarray(start, dtype=np.uint64)), _p(out))
    return out


def powers(base, n, first=None):
    out = np.empty((n, 4), dtype=np.uint64)
    f = fr_from_ints([1]) if first is None else np.ascontiguousarray(first, dtype=np.uint64)
    lib().orc_powers(_p(np.ascontiguousarray(base, dtype=np.uint64)), _p(f), C.c_size_t(n), _p(out))
    return out


class ChaCha20Rng:
    """rand_chacha::ChaCha{20,12}Rng as a linear stream of 32-bit keystream words (rand_core BlockRng): Fr::random
    (halo2curves from_u512 of 8 x next_u64) consumes 16 words, fill_bytes(32) eight.  Aligned bulk draws run in C++."""

    def __init__(self, seed32: bytes, rounds=20):
        self.seed = bytes(seed32)
        self.rounds = rounds          # 20 = ChaCha20Rng, 12 = StdRng (rand 0.8)
        self.word = 0                 # next unread keystream word

    @property
    def drawn(self):
        return self.word // 16

    def _words(self, count):
        out = np.empty(count, dtype=np.uint32)
        lib().orc_chacha_words(C.c_char_p(self.seed), C.c_int(self.rounds // 2), C.c_uint64(self.word), C.c_size_t(count), _p(out))
        self.word += count
        return out

    def fill_bytes(self, nbytes):
        """RngCore::fill_bytes for a multiple of 4 bytes (whole words are consumed)"""
        assert nbytes % 4 == 0
        return self._words(nbytes // 4).tobytes()

    def fr_random_bulk(self, n):
        """n draws as an (n, 4) Montgomery array"""
        if self.word % 16 == 0:
            out = np.empty((n, 4), dtype=np.uint64)
            lib().orc_chacha_fr_random(C.c_char_p(self.seed), C.c_int(self.rounds // 2), C.c_uint64(self.word // 16), C.c_size_t(n), _p(out))
            self.word += 16 * n
            return out
        w = self._words(16 * n).astype(object).reshape(n, 16)
        return fr_from_ints([sum(int(row[j]) << (32 * j) for j in range(16)) % R_MOD for row in w])

    def fr_random(self):
        """one draw as a canonical int"""
        return fr_to_ints(self.fr_random_bulk(1))[0]
