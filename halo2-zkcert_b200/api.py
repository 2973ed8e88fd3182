"""ctypes binding + host-side mirror of the reference operator surface.

Host field elements are numpy uint64 arrays of shape (n, 4) — the bytes of halo2curves `Fr`/`Fq`
(little-endian Montgomery limbs); G1Affine is (n, 8); G1 (Jacobian) is (n, 12).  Device-resident
data are torch CUDA tensors of dtype int64 with the same trailing shape (torch is only the
allocator / stream provider here).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzkcert_cuda.so")
_lib = None

STATUS = {0: "OK", 1: "BAD_ARG", 2: "CUDA", 3: "OOM", 10: "InvalidInstances", 11: "ConstraintSystemFailure",
          12: "BoundsFailure", 13: "Opening", 14: "Synthesis", 15: "NotEnoughRowsAvailable", 16: "Transcript"}


class ZkcError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("zkcert_cuda error %s (%d): %s" % (STATUS.get(code, "?"), code, msg))
        self.code = code


class DomainInfo(C.Structure):
    _fields_ = [("k", C.c_uint32), ("extended_k", C.c_uint32), ("j", C.c_uint32)] + \
        [(n, C.c_uint64 * 4) for n in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                                       "ifft_divisor", "extended_ifft_divisor")]


def lib_path():
    return _LIB_PATH


def lib():
    """Load libzkcert_cuda.so (building it first if the sources are newer).  Never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            from . import build as _b
            _b.build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.zkc_last_error.restype = C.c_char_p
        _lib.zkc_version.restype = C.c_char_p
        _lib.zkc_ctx_launch_count.restype = C.c_uint64
        _lib.zkc_ctx_destroy.restype = None
        _lib.zkc_domain_free.restype = None
        _lib.zkc_prove_end.restype = None
        _lib.zkc_prove_end.argtypes = [C.c_void_p]
        if hasattr(_lib, "zkc_srs_free"):
            _lib.zkc_srs_free.restype = None
    return _lib


def _np(a, cols):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError("expected array of shape (n, %d), got %r" % (cols, a.shape))
    return a


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


def _dp(t):
    """device pointer of a torch CUDA tensor"""
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


class Context:
    """One per GPU (zkc_ctx).  `use_torch_stream()` makes kernels run on torch's current stream so
    torch.cuda.Event timing brackets them."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        st = lib().zkc_ctx_create(C.c_int(device), C.byref(self._h))
        if st != 0:
            raise ZkcError(st, "zkc_ctx_create failed: no CUDA device %d (this library has no CPU fallback)" % device)
        self.device = device

    def check(self, st):
        if st != 0:
            raise ZkcError(st, lib().zkc_last_error(self._h).decode())

    def use_torch_stream(self):
        import torch
        self.check(lib().zkc_ctx_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), C.c_int(1)))

    def set_overlap(self, on=True):
        self.check(lib().zkc_ctx_set_overlap(self._h, C.c_int(1 if on else 0)))

    def sync(self):
        self.check(lib().zkc_ctx_sync(self._h))

    def set_tunable(self, name, value):
        """debug / sweep override on this ctx (zkc_ctx_set_tunable): msm_c, msm_c_pre, msm_T, ntt_two_pass_max,
        stage_min_bytes (-1 = default), team_poison, team_commit_by_column, msm_accum_occ"""
        self.check(lib().zkc_ctx_set_tunable(self._h, name.encode(), C.c_int64(int(value))))

    @property
    def launches(self):
        return int(lib().zkc_ctx_launch_count(self._h))

    def profile_enable(self, on=True):
        self.check(lib().zkc_profile_enable(self._h, C.c_int(1 if on else 0)))

    def profile_report(self):
        """{phase: {"ms": total CUDA-event ms, "n": launches}} since the last report (synchronises)."""
        import json
        buf = C.create_string_buffer(1 << 16)
        self.check(lib().zkc_profile_report(self._h, buf, C.c_size_t(len(buf))))
        return json.loads(buf.value.decode() or "{}")

    def profile_timeline(self):
        import json
        buf = C.create_string_buffer(1 << 20)
        self.check(lib().zkc_profile_timeline(self._h, buf, C.c_size_t(len(buf))))
        return json.loads(buf.value.decode() or "[]")

    # ---- team proving (zkc_team_*): one create_proof over the GPUs of a node ---------------------------------
    def team_init(self, group=None):
        """Join the NCCL team spanning the torch.distributed group (one process per GPU).  The NCCL id is
        made on rank 0 and handed round with the group's own broadcast (torch.distributed is the plumbing)."""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            self.check(lib().zkc_team_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self.check(lib().zkc_team_init(self._h, C.c_int(rank), C.c_int(world), ident))
        return rank, world

    def team_emulate(self, world):
        """all `world` shards of every partitioned step on this one GPU, collectives elided (tests)"""
        self.check(lib().zkc_team_emulate(self._h, C.c_int(world)))

    def team_leave(self):
        self.check(lib().zkc_team_leave(self._h))

    def team_info(self):
        r, w, e = C.c_int(), C.c_int(), C.c_int()
        self.check(lib().zkc_team_info(self._h, C.byref(r), C.byref(w), C.byref(e)))
        return r.value, w.value, bool(e.value)

    def close(self):
        if self._h:
            lib().zkc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- field vectors -------------------------------------------------------------------
    def field_vec_op_dev(self, field, op, a, b, out):
        ops = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "from_canonical": 4, "to_canonical": 5, "neg": 6, "square": 7}
        n = a.shape[0]
        self.check(lib().zkc_field_vec_op_dev(self._h, C.c_int(0 if field == "fr" else 1), C.c_int(ops[op]), _dp(a),
                                              None if b is None else _dp(b), _dp(out), C.c_size_t(n)))
        return out

    def batch_invert_assigned_dev(self, num, den, out):
        """poly::batch_invert_assigned on device columns: out = num / den (den = 0 -> 0)"""
        self.check(lib().zkc_batch_invert_assigned_dev(self._h, _dp(num), _dp(den), _dp(out), C.c_size_t(num.shape[0])))
        return out

    def powers_dev(self, out, base, first):
        """out[i] = first * base^i on the device (out: torch (n, 4) int64)"""
        self.check(lib().zkc_fr_powers_dev(self._h, _dp(out), C.c_size_t(out.shape[0]), _hp(_np(base, 4)), _hp(_np(first, 4))))
        return out

    # ---- best_fft ------------------------------------------------------------------------------
    def fft(self, a, omega, log_n):
        """best_fft(a, omega, log_n) on a host array; returns the transformed copy."""
        a = np.array(_np(a, 4), copy=True)
        omega = _np(omega, 4)
        assert a.shape[0] == 1 << log_n
        self.check(lib().zkc_fft_fr(self._h, _hp(a), _hp(omega), C.c_uint32(log_n)))
        return a

    def fft_dev(self, a_dev, omega, log_n, ncols=1):
        omega = _np(omega, 4)
        self.check(lib().zkc_fft_fr_dev(self._h, _dp(a_dev), _hp(omega), C.c_uint32(log_n), C.c_uint32(ncols)))
        return a_dev

    # ---- best_multiexp -------------------------------------------------------------------------
    def msm(self, scalars, bases):
        scalars, bases = _np(scalars, 4), _np(bases, 8)
        assert scalars.shape[0] == bases.shape[0]
        out = np.zeros((1, 12), dtype=np.uint64)
        self.check(lib().zkc_msm_g1(self._h, _hp(scalars), _hp(bases), C.c_size_t(scalars.shape[0]), _hp(out)))
        return out

    def msm_dev(self, scalars_dev, bases_dev, n, ncols=1):
        out = np.zeros((ncols, 12), dtype=np.uint64)
        self.check(lib().zkc_msm_g1_dev(self._h, _dp(scalars_dev), _dp(bases_dev), C.c_size_t(n), C.c_uint32(ncols), _hp(out)))
        return out


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def best_fft(a, omega, log_n, ctx=None):
    """halo2_proofs::arithmetic::best_fft — natural order in/out; returns a new array."""
    return (ctx or default_context()).fft(a, omega, log_n)


def best_multiexp(coeffs, bases, ctx=None):
    """halo2_proofs::arithmetic::best_multiexp — returns the Jacobian result (1, 12), normalised."""
    return (ctx or default_context()).msm(coeffs, bases)


def jac_to_affine(j):
    """normalised Jacobian (z = 1 or 0) -> affine (x, y); identity -> (0, 0)"""
    j = np.asarray(j, dtype=np.uint64).reshape(-1, 12)
    out = j[:, :8].copy()
    ident = ~j[:, 8:].any(axis=1)
    out[ident] = 0
    return out


class EvaluationDomain:
    """halo2_proofs::poly::EvaluationDomain::new(j, k) and its conversions."""

    def __init__(self, j, k, ctx=None, zeta_choice=0):
        self.ctx = ctx or default_context()
        self._h = C.c_void_p()
        self.ctx.check(lib().zkc_domain_create(self.ctx._h, C.c_uint32(j), C.c_uint32(k), C.c_int(zeta_choice), C.byref(self._h)))
        info = DomainInfo()
        self.ctx.check(lib().zkc_domain_get_info(self._h, C.byref(info)))
        self.k, self.extended_k, self.j = info.k, info.extended_k, info.j
        self.n, self.extended_n = 1 << info.k, 1 << info.extended_k
        for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv", "ifft_divisor",
                     "extended_ifft_divisor"):
            setattr(self, name, np.array([list(getattr(info, name))], dtype=np.uint64))

    def __del__(self):
        try:
            if self._h:
                lib().zkc_domain_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def get_omega(self):
        return self.omega

    def get_extended_omega(self):
        return self.extended_omega

    def lagrange_to_coeff(self, a):
        a = np.array(_np(a, 4), copy=True)
        assert a.shape[0] == self.n
        self.ctx.check(lib().zkc_lagrange_to_coeff(self.ctx._h, self._h, _hp(a)))
        return a

    def coeff_to_lagrange(self, a):
        a = np.array(_np(a, 4), copy=True)
        assert a.shape[0] == self.n
        self.ctx.check(lib().zkc_coeff_to_lagrange(self.ctx._h, self._h, _hp(a)))
        return a

    def coeff_to_extended(self, a):
        a = _np(a, 4)
        assert a.shape[0] == self.n
        out = np.empty((self.extended_n, 4), dtype=np.uint64)
        self.ctx.check(lib().zkc_coeff_to_extended(self.ctx._h, self._h, _hp(a), _hp(out)))
        return out

    def extended_to_coeff(self, a):
        a = np.array(_np(a, 4), copy=True)
        assert a.shape[0] == self.extended_n
        self.ctx.check(lib().zkc_extended_to_coeff(self.ctx._h, self._h, _hp(a)))
        return a

    # device-resident, batched
    def lagrange_to_coeff_dev(self, t, ncols=1):
        self.ctx.check(lib().zkc_lagrange_to_coeff_dev(self.ctx._h, self._h, _dp(t), C.c_uint32(ncols)))
        return t

    def coeff_to_lagrange_dev(self, t, ncols=1):
        self.ctx.check(lib().zkc_coeff_to_lagrange_dev(self.ctx._h, self._h, _dp(t), C.c_uint32(ncols)))
        return t

    def coeff_to_extended_dev(self, src, dst, ncols=1):
        self.ctx.check(lib().zkc_coeff_to_extended_dev(self.ctx._h, self._h, _dp(src), _dp(dst), C.c_uint32(ncols)))
        return dst

    def coeff_to_extended_classes_dev(self, src, dst, ncols, c0, c1):
        """classes [c0, c1) of the extended coset, class-major (the residue-class split of team proving)"""
        self.ctx.check(lib().zkc_coeff_to_extended_classes_dev(self.ctx._h, self._h, _dp(src), _dp(dst), C.c_uint32(ncols), C.c_uint32(c0), C.c_uint32(c1)))
        return dst

    def extended_classes_to_natural_dev(self, src, dst):
        self.ctx.check(lib().zkc_extended_classes_to_natural_dev(self.ctx._h, self._h, _dp(src), _dp(dst)))
        return dst

    def extended_to_coeff_dev(self, t, ncols=1):
        self.ctx.check(lib().zkc_extended_to_coeff_dev(self.ctx._h, self._h, _dp(t), C.c_uint32(ncols)))
        return t

    def divide_by_vanishing_poly_dev(self, t):
        self.ctx.check(lib().zkc_divide_by_vanishing_dev(self.ctx._h, self._h, _dp(t)))
        return t


class ParamsKZG:
    """halo2_proofs::poly::kzg::commitment::ParamsKZG<Bn256> with g / g_lagrange resident in HBM."""

    def __init__(self, k, g=None, g_lagrange=None, s=None, ctx=None):
        self.ctx = ctx or default_context()
        self.k, self.n = k, 1 << k
        self._h = C.c_void_p()
        if g is not None:
            g, gl = _np(g, 8), _np(g_lagrange, 8)
            assert g.shape[0] == self.n and gl.shape[0] == self.n
            self.ctx.check(lib().zkc_srs_load(self.ctx._h, C.c_uint32(k), _hp(g), _hp(gl), C.byref(self._h)))
        else:
            s = _np(s, 4)
            self.ctx.check(lib().zkc_srs_setup(self.ctx._h, C.c_uint32(k), _hp(s), C.byref(self._h)))

    @classmethod
    def setup(cls, k, s, ctx=None):
        """ParamsKZG::setup(k, rng) with the secret s = Fr::random(rng) supplied by the caller."""
        return cls(k, s=s, ctx=ctx)

    def __del__(self):
        try:
            if self._h:
                lib().zkc_srs_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def write(self, path, s_g2, g2=None):
        """ParamsKZG::write: kzg_bn254_{k}.srs in RawBytes form (g, g_lagrange from the device; g2 / s_g2 live on the host)"""
        data = params_to_bytes(self.k, self.get_g(0), self.get_g(1), g2_generator() if g2 is None else g2, s_g2)
        with open(path, "wb") as f:
            f.write(data)
        return len(data)

    @classmethod
    def read(cls, path, ctx=None, checked=True):
        """ParamsKZG::read: loads the file and uploads g / g_lagrange; returns (params, g2, s_g2)"""
        with open(path, "rb") as f:
            k, g, gl, g2, s_g2 = params_from_bytes(f.read(), checked)
        return cls(k, g=g, g_lagrange=gl, ctx=ctx), g2, s_g2

    def get_g(self, basis=0):
        out = np.empty((self.n, 8), dtype=np.uint64)
        self.ctx.check(lib().zkc_srs_get(self.ctx._h, self._h, C.c_int(basis), _hp(out)))
        return out

    def commit(self, poly):
        poly = _np(poly, 4)
        out = np.zeros((1, 12), dtype=np.uint64)
        self.ctx.check(lib().zkc_commit(self.ctx._h, self._h, C.c_int(0), _hp(poly), C.c_size_t(poly.shape[0]), _hp(out)))
        return out

    def commit_lagrange(self, poly):
        poly = _np(poly, 4)
        out = np.zeros((1, 12), dtype=np.uint64)
        self.ctx.check(lib().zkc_commit(self.ctx._h, self._h, C.c_int(1), _hp(poly), C.c_size_t(poly.shape[0]), _hp(out)))
        return out

    def commit_dev(self, polys_dev, length, ncols=1, basis=0):
        out = np.zeros((ncols, 12), dtype=np.uint64)
        self.ctx.check(lib().zkc_commit_dev(self.ctx._h, self._h, C.c_int(basis), _dp(polys_dev), C.c_size_t(length), C.c_uint32(ncols), _hp(out)))
        return out


# ---- ProvingKey / create_proof --------------------------------------------------------------------
class ProveOpts(C.Structure):
    _fields_ = [("transcript", C.c_int), ("multiopen", C.c_int), ("advice_blinding", C.c_int), ("blind_draws", C.c_int),
                ("point_format", C.c_int), ("rng_kind", C.c_int), ("rng_seed", C.c_uint8 * 32), ("lookup_fill", C.c_int),
                ("random_poly", C.c_int), ("random_poly_threads", C.c_uint32)]


class RandomPoly(C.Structure):
    """zkc_random_poly: the vanishing argument's random polynomial as the step API takes it"""
    _fields_ = [("kind", C.c_int), ("scalars", C.c_void_p), ("seed", C.c_uint8 * 32), ("rng_kind", C.c_int), ("first_word", C.c_uint64),
                ("seeds", C.c_void_p), ("nseeds", C.c_uint32), ("chunk_len", C.c_uint64)]


LOOKUP_FILL = {"pse": 0, "axiom": 1}
RANDOM_POLY = {"serial": 0, "chunked": 1}


def _prove_opts(transcript="blake2b", multiopen="shplonk", advice_blinding="axiom", blind_draws=False, point_format=0, rng="chacha20",
                rng_seed=bytes(32), lookup_fill="pse", random_poly="serial", random_poly_threads=1):
    """zkc_prove_opts from the reference's vocabulary (transcript / multiopen names, OPEN switches of SURVEY 8c)"""
    o = ProveOpts()
    o.transcript = {"blake2b": 0, "keccak": 1, "evm": 2, "poseidon": 3}[transcript]
    o.multiopen = {"shplonk": 0, "gwc": 1}[multiopen]
    o.advice_blinding = {"axiom": 0, "pse": 1}[advice_blinding]
    o.blind_draws = 1 if blind_draws else 0
    o.point_format = point_format
    o.rng_kind = {"chacha20": 0, "std": 1, "chacha12": 1}[rng]
    o.rng_seed[:] = list(rng_seed)
    o.lookup_fill = LOOKUP_FILL[lookup_fill]
    o.random_poly = RANDOM_POLY[random_poly]
    o.random_poly_threads = int(random_poly_threads)
    return o


def _instances(cs, instances):
    """ctypes views of the instance columns; upstream's Error::InvalidInstances when their number is not cs.num_instance_columns"""
    inst = [_np(i, 4) if len(i) else np.zeros((0, 4), dtype=np.uint64) for i in instances]
    ptrs = (C.c_void_p * max(len(inst), 1))(*[i.ctypes.data for i in inst])
    lens = (C.c_size_t * max(len(inst), 1))(*[i.shape[0] for i in inst])
    return inst, ptrs, lens, C.c_size_t(len(inst))


class ProvingKey:
    """halo2_proofs::plonk::ProvingKey resident on the device (zkc_pk).

    cs: circuit.ConstraintSystem; fixed / sigma: (num_columns * n, 4) Montgomery Lagrange columns
    (pk.fixed_values, pk.permutation.permutations); transcript_repr: (1, 4) Montgomery."""

    def __init__(self, params, cs, fixed, sigma, transcript_repr, zeta_choice=0, copies=None, file_bytes=None, num_selectors=0,
                 file_format="unchecked"):
        """keygen_pk from the Lagrange columns: fixed + sigma (zkc_pk_load), fixed + copy constraints (`copies`: (m, 4) uint32
        rows (left column, left row, right column, right row) over cs.permutation's columns; zkc_keygen_pk assembles the
        permutation), or the bytes of a `.pk` file (zkc_pk_read; file_format "raw" = SerdeFormat::RawBytes, checked)."""
        self.params, self.cs, self.ctx = params, cs, params.ctx
        n = cs.n
        blob = cs.serialize()
        tr = _np(transcript_repr, 4)
        self._h = C.c_void_p()
        if file_bytes is not None:
            raw = np.frombuffer(file_bytes, dtype=np.uint8)
            self.ctx.check(lib().zkc_pk_read(self.ctx._h, params._h, blob, C.c_size_t(len(blob)), _hp(raw), C.c_size_t(raw.size), C.c_uint32(num_selectors),
                                             C.c_int({"raw": 0, "unchecked": 1}[file_format]), _hp(tr), C.c_int(zeta_choice), C.byref(self._h)))
        else:
            fixed = _np(fixed, 4) if cs.num_fixed else np.zeros((0, 4), dtype=np.uint64)
            assert fixed.shape[0] == cs.num_fixed * n
            if copies is not None:
                cp = np.ascontiguousarray(np.asarray(copies, dtype=np.uint32).reshape(-1, 4))
                self.ctx.check(lib().zkc_keygen_pk(self.ctx._h, params._h, blob, C.c_size_t(len(blob)), _hp(fixed), _hp(cp), C.c_size_t(cp.shape[0]),
                                                   _hp(tr), C.c_int(zeta_choice), C.byref(self._h)))
            else:
                sigma = _np(sigma, 4) if cs.permutation else np.zeros((0, 4), dtype=np.uint64)
                assert sigma.shape[0] == len(cs.permutation) * n
                self.ctx.check(lib().zkc_pk_load(self.ctx._h, params._h, blob, C.c_size_t(len(blob)), _hp(fixed), _hp(sigma), _hp(tr),
                                                 C.c_int(zeta_choice), C.byref(self._h)))
        info = (C.c_uint32 * 8)()
        lib().zkc_pk_info(self._h, info)
        self.k, self.extended_k, self.degree, self.blinding_factors, self.num_sets, self.num_lookups = list(info)[:6]

    @classmethod
    def keygen(cls, params, cs, fixed, copies, transcript_repr, zeta_choice=0):
        """keygen_pk as gen_pk reaches it: fixed columns + the copy constraints synthesis recorded"""
        return cls(params, cs, fixed, None, transcript_repr, zeta_choice, copies=copies)

    @classmethod
    def read(cls, params, cs, data, transcript_repr, num_selectors=0, file_format="unchecked", zeta_choice=0):
        """ProvingKey::read (`*.pk`, SerdeFormat::RawBytesUnchecked by default, as read_pk uses)"""
        return cls(params, cs, None, None, transcript_repr, zeta_choice, file_bytes=data, num_selectors=num_selectors, file_format=file_format)

    def write(self, selectors=None, be=True):
        """ProvingKey::write -> bytes.  selectors: (num_selectors, ceil(n / 8)) uint8 bit-packed activations (vk.selectors) or None"""
        ns = 0 if selectors is None else int(selectors.shape[0])
        lib().zkc_pk_file_size.restype = C.c_size_t
        size = lib().zkc_pk_file_size(self._h, C.c_uint32(ns))
        out = np.zeros(size, dtype=np.uint8)
        sel = None if selectors is None else _hp(np.ascontiguousarray(selectors, dtype=np.uint8))
        self.ctx.check(lib().zkc_pk_write(self.ctx._h, self._h, sel, C.c_uint32(ns), C.c_int(1 if be else 0), _hp(out), C.c_size_t(size)))
        return out.tobytes()

    def sigma(self):
        """pk.permutation.permutations: (n_perm_columns * n, 4) Montgomery Lagrange columns"""
        out = np.zeros((len(self.cs.permutation) * self.cs.n, 4), dtype=np.uint64)
        self.ctx.check(lib().zkc_pk_get_sigma(self.ctx._h, self._h, _hp(out)))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().zkc_pk_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def commitments(self):
        """(vk.fixed_commitments, vk.permutation.commitments) as affine (m, 8) arrays"""
        f = np.zeros((self.cs.num_fixed, 8), dtype=np.uint64)
        s = np.zeros((len(self.cs.permutation), 8), dtype=np.uint64)
        self.ctx.check(lib().zkc_pk_get_commitments(self.ctx._h, self._h, _hp(f), _hp(s)))
        return f, s


def _advice_ptr(pk, advice):
    on_dev = hasattr(advice, "is_cuda")
    if on_dev:
        if advice.shape[0] != pk.cs.num_advice * pk.cs.n:
            raise ValueError("advice: expected num_advice * n rows")
        return advice, _dp(advice), 1
    advice = _np(advice, 4)
    if advice.shape[0] != pk.cs.num_advice * pk.cs.n:
        raise ValueError("advice: expected num_advice * n rows")
    return advice, _hp(advice), 0


def create_proof(pk, advice, instances, rng_seed, transcript="blake2b", multiopen="shplonk", advice_blinding="axiom", blind_draws=False,
                 point_format=0, rng="chacha20", lookup_fill="pse", random_poly="serial", random_poly_threads=1, ctx=None):
    """plonk::create_proof for one circuit.  advice: (num_advice * n, 4) Montgomery, numpy (host) or
    torch CUDA tensor (already resident); instances: list of (len, 4) Montgomery arrays; rng_seed: 32
    bytes for ChaCha20Rng::from_seed.  Returns the proof bytes.  `ctx`: run on another zkc_ctx of the SAME device than the one
    the key was loaded through (its own streams and scratch; the resident SRS / key are only read) — several proofs in flight."""
    ctx = ctx or pk.ctx
    o = _prove_opts(transcript, multiopen, advice_blinding, blind_draws, point_format, rng, rng_seed, lookup_fill, random_poly, random_poly_threads)
    advice, adv_ptr, on_dev = _advice_ptr(pk, advice)
    inst, ptrs, lens, ninst = _instances(pk.cs, instances)
    cap = 1 << 20
    buf = (C.c_uint8 * cap)()
    plen = C.c_size_t(0)
    ctx.check(lib().zkc_prove(ctx._h, pk._h, adv_ptr, C.c_int(on_dev), ptrs, lens, ninst, C.byref(o), buf, C.c_size_t(cap), C.byref(plen)))
    return bytes(buf[:plen.value])


class ProverSession:
    """The step API of create_proof (zkc_prove_begin ... zkc_prove_end): the caller keeps the Fiat-Shamir transcript and the
    RNG, the library does the O(n) work of each round.  Scalars go in and out as (m, 4) Montgomery arrays, points come back as
    (m, 8) affine arrays (identity = zeros)."""

    def __init__(self, pk, advice, instances, advice_tails=None, early_random=None):
        self.pk, self.ctx = pk, pk.ctx
        cs = pk.cs
        advice, adv_ptr, on_dev = _advice_ptr(pk, advice)
        inst, ptrs, lens, ninst = _instances(cs, instances)
        tails = None if advice_tails is None else _np(advice_tails, 4)
        self._keep = [advice, inst, tails, early_random]
        self._h = C.c_void_p()
        out = np.zeros((cs.num_advice, 8), dtype=np.uint64)
        self.ctx.check(lib().zkc_prove_begin(self.ctx._h, pk._h, adv_ptr, C.c_int(on_dev), ptrs, lens, ninst, None if tails is None else _hp(tails),
                                             None if early_random is None else C.byref(early_random._c), C.byref(self._h), _hp(out)))
        self.advice_commitments = out

    def lookups(self, theta, tails, lookup_fill="pse"):
        out = np.zeros((2 * self.pk.num_lookups, 8), dtype=np.uint64)
        tails = _np(tails, 4) if self.pk.num_lookups else np.zeros((1, 4), dtype=np.uint64)
        self.ctx.check(lib().zkc_prove_lookups(self._h, _hp(_np(theta, 4)), _hp(tails), C.c_int(LOOKUP_FILL[lookup_fill]), _hp(out)))
        return out

    def products(self, beta, gamma, tails):
        m = self.pk.num_sets + self.pk.num_lookups
        out = np.zeros((m, 8), dtype=np.uint64)
        tails = _np(tails, 4) if m else np.zeros((1, 4), dtype=np.uint64)
        self.ctx.check(lib().zkc_prove_products(self._h, _hp(_np(beta, 4)), _hp(_np(gamma, 4)), _hp(tails), _hp(out)))
        return out

    def vanishing(self, random=None):
        out = np.zeros((1, 8), dtype=np.uint64)
        self._keep.append(random)
        self.ctx.check(lib().zkc_prove_vanishing(self._h, None if random is None else C.byref(random._c), _hp(out)))
        return out

    def quotient(self, y):
        out = np.zeros((self.pk.degree - 1, 8), dtype=np.uint64)
        self.ctx.check(lib().zkc_prove_quotient(self._h, _hp(_np(y, 4)), _hp(out)))
        return out

    def evals(self, x):
        cap = 1 << 14
        out = np.zeros((cap, 4), dtype=np.uint64)
        cnt = C.c_size_t(0)
        self.ctx.check(lib().zkc_prove_evals(self._h, _hp(_np(x, 4)), _hp(out), C.c_size_t(cap), C.byref(cnt)))
        return out[:cnt.value].copy()

    def open_shplonk_h(self, y, v):
        out = np.zeros((1, 8), dtype=np.uint64)
        self.ctx.check(lib().zkc_prove_open_shplonk_h(self._h, _hp(_np(y, 4)), _hp(_np(v, 4)), _hp(out)))
        return out

    def open_shplonk_w(self, u):
        out = np.zeros((1, 8), dtype=np.uint64)
        self.ctx.check(lib().zkc_prove_open_shplonk_w(self._h, _hp(_np(u, 4)), _hp(out)))
        return out

    def open_gwc(self, v):
        cap = 256
        out = np.zeros((cap, 8), dtype=np.uint64)
        cnt = C.c_size_t(0)
        self.ctx.check(lib().zkc_prove_open_gwc(self._h, _hp(_np(v, 4)), _hp(out), C.c_size_t(cap), C.byref(cnt)))
        return out[:cnt.value].copy()

    def end(self):
        if self._h:
            lib().zkc_prove_end(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.end()
        except Exception:
            pass


class RandomPolySpec:
    """zkc_random_poly builder: scalars (n, 4) | keystream (seed, rng, first_word) | chunk seeds (list of 32-byte seeds, chunk_len)"""

    def __init__(self, scalars=None, seed=None, rng="chacha20", first_word=0, seeds=None, chunk_len=0):
        c = RandomPoly()
        if scalars is not None:
            self._a = _np(scalars, 4)
            c.kind, c.scalars = 0, self._a.ctypes.data
        elif seeds is not None:
            self._a = np.frombuffer(b"".join(seeds), dtype=np.uint8).copy()
            c.kind, c.seeds, c.nseeds, c.chunk_len = 2, self._a.ctypes.data, len(seeds), int(chunk_len)
        else:
            c.kind, c.rng_kind, c.first_word = 1, (0 if rng == "chacha20" else 1), int(first_word)
            c.seed[:] = list(seed)
        self._c = c


# ---- keygen helpers (host only) -------------------------------------------------------------------------------------
def keygen_permutation_mapping(k, num_columns, copies):
    """permutation::keygen::Assembly over the copy list ((m, 4): left col, left row, right col, right row): flat mapping
    [col * n + row] -> col' * n + row' (uint64)"""
    cp = np.ascontiguousarray(np.asarray(copies, dtype=np.uint32).reshape(-1, 4))
    out = np.zeros(num_columns << k, dtype=np.uint64)
    st = lib().zkc_keygen_permutation_mapping(C.c_uint32(k), C.c_uint32(num_columns), _hp(cp), C.c_size_t(cp.shape[0]), _hp(out))
    if st != 0:
        raise ZkcError(st, "zkc_keygen_permutation_mapping")
    return out


def compress_selectors(k, activations, max_degrees, max_degree):
    """ConstraintSystem::compress_selectors.  activations: (num_selectors, n) 0/1; returns (combination_of, root_of,
    combination_len, columns) with columns (num_combinations, n) uint32"""
    act = np.ascontiguousarray(activations, dtype=np.uint8)
    S, n = act.shape
    assert n == 1 << k
    md = np.ascontiguousarray(max_degrees, dtype=np.uint32)
    comb, root, clen = (np.zeros(S, dtype=np.uint32) for _ in range(3))
    cols = np.zeros((max(S, 1), n), dtype=np.uint32)
    ncomb = C.c_uint32(0)
    st = lib().zkc_keygen_compress_selectors(C.c_uint32(k), C.c_uint32(S), _hp(act), _hp(md), C.c_uint32(max_degree), _hp(comb), _hp(root),
                                             _hp(clen), _hp(cols), C.byref(ncomb))
    if st != 0:
        raise ZkcError(st, "zkc_keygen_compress_selectors")
    return comb, root, clen, cols[:ncomb.value].copy()


class PkFileLayout(C.Structure):
    _fields_ = [(nm, C.c_uint64) for nm in ("k_off", "num_fixed_off", "fixed_commitments_off", "perm_commitments_off", "selectors_off", "l0_off",
                                            "l_last_off", "l_active_row_off", "fixed_values_off", "fixed_polys_off", "fixed_cosets_off",
                                            "perm_values_off", "perm_polys_off", "perm_cosets_off", "total")]


def pk_file_layout(k, extended_k, num_fixed, num_perm, num_selectors=0):
    L = PkFileLayout()
    st = lib().zkc_pk_file_layout(C.c_uint32(k), C.c_uint32(extended_k), C.c_uint32(num_fixed), C.c_uint32(num_perm), C.c_uint32(num_selectors), C.byref(L))
    if st != 0:
        raise ZkcError(st, "zkc_pk_file_layout")
    return L


# ---- ParamsKZG files (host only) -----------------------------------------------------------------------------------
def params_to_bytes(k, g, g_lagrange, g2, s_g2):
    """ParamsKZG::write (SerdeFormat::RawBytes): the bytes of a kzg_bn254_{k}.srs file"""
    lib().zkc_params_size.restype = C.c_size_t
    size = lib().zkc_params_size(C.c_uint32(k))
    buf = np.zeros(size, dtype=np.uint8)
    st = lib().zkc_params_write(C.c_uint32(k), _hp(_np(g, 8)), _hp(_np(g_lagrange, 8)), _hp(_np(g2, 16)), _hp(_np(s_g2, 16)), _hp(buf), C.c_size_t(size))
    if st != 0:
        raise ZkcError(st, "zkc_params_write")
    return buf.tobytes()


def params_from_bytes(data, checked=True):
    """ParamsKZG::read: (k, g, g_lagrange, g2, s_g2) as numpy arrays; checked=True validates every point (RawBytes)"""
    raw = np.frombuffer(data, dtype=np.uint8)
    k = C.c_uint32(0)
    st = lib().zkc_params_read(_hp(raw), C.c_size_t(raw.size), C.c_int(0), C.byref(k), None, None, None, None)
    if st != 0:
        raise ZkcError(st, "zkc_params_read: not a ParamsKZG RawBytes file (size does not match k)")
    n = 1 << k.value
    g, gl = np.zeros((n, 8), dtype=np.uint64), np.zeros((n, 8), dtype=np.uint64)
    g2, s_g2 = np.zeros((1, 16), dtype=np.uint64), np.zeros((1, 16), dtype=np.uint64)
    st = lib().zkc_params_read(_hp(raw), C.c_size_t(raw.size), C.c_int(1 if checked else 0), C.byref(k), _hp(g), _hp(gl), _hp(g2), _hp(s_g2))
    if st != 0:
        raise ZkcError(st, "zkc_params_read: invalid point in the file")
    return k.value, g, gl, g2, s_g2


# ---- verify_proof (host only) --------------------------------------------------------------------------------------
def g2_generator():
    """(1, 16) uint64: halo2curves G2Affine::generator() (x.c0, x.c1, y.c0, y.c1; Montgomery limbs)"""
    out = np.zeros((1, 16), dtype=np.uint64)
    assert lib().zkc_g2_generator(_hp(out)) == 0
    return out


def g2_mul(point, scalar):
    """[scalar]P on G2 (scalar: (1, 4) Montgomery Fr) — ParamsKZG::setup's s_g2 = [s]G2"""
    out = np.zeros((1, 16), dtype=np.uint64)
    st = lib().zkc_g2_mul(_hp(_np(point, 16)), _hp(_np(scalar, 4)), _hp(out))
    if st != 0:
        raise ZkcError(st, "zkc_g2_mul")
    return out


def pairing_check(g1s, g2s):
    """prod_i e(g1s[i], g2s[i]) == 1 ; g1s (m, 8) affine, g2s (m, 16)"""
    g1s, g2s = _np(g1s, 8), _np(g2s, 16)
    assert g1s.shape[0] == g2s.shape[0]
    ok = C.c_int(0)
    st = lib().zkc_pairing_check(_hp(g1s), _hp(g2s), C.c_size_t(g1s.shape[0]), C.byref(ok))
    if st != 0:
        raise ZkcError(st, "zkc_pairing_check: point not on the curve")
    return bool(ok.value)


def verify_proof(cs, fixed_commitments, sigma_commitments, transcript_repr, g1_gen, g2, s_g2, instances, proof,
                 transcript="blake2b", multiopen="shplonk", point_format=0):
    """plonk::verify_proof (zkc_verify, host only).  cs: circuit.ConstraintSystem; commitments: (m, 8) affine Montgomery
    arrays (ProvingKey.commitments()); transcript_repr (1, 4); g1_gen (1, 8) = params.g[0]; g2 / s_g2 (1, 16);
    instances: list of (len, 4) Montgomery arrays.  Returns True iff the proof is accepted."""
    blob = cs.serialize()
    f = _np(fixed_commitments, 8) if cs.num_fixed else np.zeros((0, 8), dtype=np.uint64)
    sg = _np(sigma_commitments, 8) if cs.permutation else np.zeros((0, 8), dtype=np.uint64)
    inst, ptrs, lens, ninst = _instances(cs, instances)
    o = _prove_opts(transcript, multiopen, point_format=point_format)
    buf = (C.c_uint8 * max(len(proof), 1)).from_buffer_copy(bytes(proof) or b"\0")
    ok = C.c_int(0)
    st = lib().zkc_verify(blob, C.c_size_t(len(blob)), _hp(f), _hp(sg), _hp(_np(transcript_repr, 4)), _hp(_np(g1_gen, 8)), _hp(_np(g2, 16)),
                          _hp(_np(s_g2, 16)), ptrs, lens, ninst, buf, C.c_size_t(len(proof)), C.byref(o), C.byref(ok))
    if st != 0:
        raise ZkcError(st, "zkc_verify: instances.len() != cs.num_instance_columns" if st == 10 else "zkc_verify")
    return bool(ok.value)


class AdviceColumn(C.Structure):
    _fields_ = [("kind", C.c_int), ("data", C.c_void_p)]


COMPACT_KINDS = {"fr": 0, "bits": 1, "u8": 2, "u16": 3, "u64": 4}


class CompactAdvice:
    """Witness columns in compact host form for zkc_prove_compact: list of (kind, numpy array)."""

    def __init__(self, columns):
        self.columns = [(COMPACT_KINDS[k], np.ascontiguousarray(a)) for k, a in columns]
        self.nbytes = sum(a.nbytes for _, a in self.columns)

    @staticmethod
    def pack_canonical(col_limbs):
        """choose the tightest encoding for one column of canonical (n, 4) uint64 limbs"""
        hi = col_limbs[:, 1:].any()
        lo = col_limbs[:, 0]
        if hi:
            raise ValueError("column does not fit 64 bits; pass it as ('fr', montgomery limbs)")
        m = int(lo.max()) if lo.size else 0
        if m <= 1:
            return ("bits", np.packbits(lo.astype(np.uint8), bitorder="little"))
        if m < 256:
            return ("u8", lo.astype(np.uint8))
        if m < 65536:
            return ("u16", lo.astype(np.uint16))
        return ("u64", lo.astype(np.uint64))


def create_proof_compact(pk, compact, instances, rng_seed, transcript="blake2b", multiopen="shplonk", advice_blinding="axiom",
                         blind_draws=False, point_format=0, rng="chacha20", lookup_fill="pse", random_poly="serial", random_poly_threads=1):
    """create_proof with the witness handed over in compact host form (CompactAdvice)"""
    ctx = pk.ctx
    if len(compact.columns) != pk.cs.num_advice:
        raise ValueError("compact witness: expected num_advice columns")
    o = _prove_opts(transcript, multiopen, advice_blinding, blind_draws, point_format, rng, rng_seed, lookup_fill, random_poly, random_poly_threads)
    cols = (AdviceColumn * max(len(compact.columns), 1))()
    for i, (kind, arr) in enumerate(compact.columns):
        cols[i].kind = kind
        cols[i].data = arr.ctypes.data
    inst, ptrs, lens, ninst = _instances(pk.cs, instances)
    cap = 1 << 20
    buf = (C.c_uint8 * cap)()
    plen = C.c_size_t(0)
    ctx.check(lib().zkc_prove_compact(ctx._h, pk._h, cols, ptrs, lens, ninst, C.byref(o), buf, C.c_size_t(cap), C.byref(plen)))
    return bytes(buf[:plen.value])


def seed_from_u64(state):
    out = (C.c_uint8 * 32)()
    lib().zkc_seed_from_u64(C.c_uint64(state), out)
    return bytes(out)


def fr_random_stream(seed, count, skip=0, rng="chacha20"):
    """Fr::random draws of ChaCha20Rng / StdRng ::from_seed(seed) as (count, 4) Montgomery limbs (host-side helper)"""
    out = np.zeros((count, 4), dtype=np.uint64)
    s = (C.c_uint8 * 32)(*list(seed))
    st = lib().zkc_rng_fr_random(s, C.c_int(0 if rng == "chacha20" else 1), C.c_uint64(skip), _hp(out), C.c_size_t(count))
    if st != 0:
        raise ZkcError(st, "zkc_rng_fr_random")
    return out


def g1_sum(points):
    """host-side sum of normalised Jacobian points (n, 12) -> (1, 12): point-sharded MSM epilogue"""
    pts = _np(np.asarray(points, dtype=np.uint64).reshape(-1, 12), 12)
    out = np.zeros((1, 12), dtype=np.uint64)
    st = lib().zkc_g1_sum(_hp(pts), C.c_size_t(pts.shape[0]), _hp(out))
    if st != 0:
        raise ZkcError(st, "zkc_g1_sum")
    return out


def poseidon_spec():
    """(round constants [65][3], MDS [3][3]) of the Poseidon transcript as Python ints"""
    c = np.zeros((195, 4), dtype=np.uint64)
    m = np.zeros((9, 4), dtype=np.uint64)
    st = lib().zkc_poseidon_spec(_hp(c), _hp(m))
    if st != 0:
        raise ZkcError(st, "zkc_poseidon_spec")
    toint = lambda row: sum(int(row[i]) << (64 * i) for i in range(4))
    cs = [toint(r) for r in c]
    ms = [toint(r) for r in m]
    return [cs[3 * i:3 * i + 3] for i in range(65)], [ms[3 * i:3 * i + 3] for i in range(3)]
