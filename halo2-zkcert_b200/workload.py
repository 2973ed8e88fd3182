"""Builds a complete proving workload (SRS, proving key, resident witness) for a synthetic circuit
using only the product library — what bench.py, smoke() and the multi-GPU driver run.

Off the timed path: this is the work `gen_srs` + `gen_pk` + witness synthesis do in the reference
(/root/reference/src/helpers.rs:201-216, 236-266).  Montgomery conversion, the sigma table and the
SRS are produced on the device; torch is only the allocator here.
"""
import numpy as np

from . import api, synth
from .circuit import R_MOD

GEN_SRS_SEED = bytes(32)     # halo2-base gen_srs: ChaCha20Rng::from_seed([0; 32])


def to_mont_dev(ctx, ints):
    """canonical Python ints -> torch CUDA (len, 4) int64 Montgomery limbs"""
    import torch
    t = torch.from_numpy(synth.ints_to_limbs(ints).view(np.int64)).cuda(ctx.device)
    out = torch.empty_like(t)
    ctx.field_vec_op_dev("fr", "from_canonical", t, None, out)
    return out


def limbs_to_mont_dev(ctx, limbs):
    """canonical (len, 4) uint64 limbs -> torch CUDA Montgomery limbs"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(limbs).view(np.int64)).cuda(ctx.device)
    out = torch.empty_like(t)
    ctx.field_vec_op_dev("fr", "from_canonical", t, None, out)
    return out


def to_host(t):
    return t.cpu().numpy().view(np.uint64)


class Workload:
    pass


def build(ctx, k, num_gate_cols, seed=0, circ=None, params=None, shape="base"):
    import torch
    w = Workload()
    w.ctx = ctx
    if circ is None:
        gen = {"sha_bit": synth.make_sha_bit_circuit, "base_fast": synth.make_base_circuit_fast, "base": synth.make_base_circuit}[shape]
        circ = gen(k, num_gate_cols, seed=seed)
    w.circ = circ
    cs = w.circ.cs
    n = cs.n
    if params is None:
        s = api.fr_random_stream(GEN_SRS_SEED, 1)
        params = api.ParamsKZG.setup(k, s, ctx=ctx)
    w.params = params
    flat = lambda cols: [v for c in cols for v in c]
    if hasattr(w.circ, "advice_limbs"):
        fixed = limbs_to_mont_dev(ctx, np.concatenate(w.circ.fixed_limbs))
        w.advice_dev = limbs_to_mont_dev(ctx, np.concatenate(w.circ.advice_limbs))
    else:
        fixed = to_mont_dev(ctx, flat(w.circ.fixed))
        w.advice_dev = to_mont_dev(ctx, flat(w.circ.advice))
    w.instances = [to_host(to_mont_dev(ctx, c)) if len(c) else np.zeros((0, 4), dtype=np.uint64) for c in w.circ.instances]
    # keygen_pk behind the C ABI (zkc_keygen_pk): the permutation is assembled from the copy constraints on the host (C++), the
    # sigma columns DELTA^col' * omega^row' are expanded on the device, then polys / cosets / vk commitments as for any key
    tr = to_host(to_mont_dev(ctx, [w.circ.transcript_repr()]))
    w.transcript_repr = tr          # (1, 4) Montgomery: vk.transcript_repr, also what verify_proof absorbs first
    copies = np.asarray(w.circ.copies, dtype=np.uint32).reshape(-1, 4)
    w.pk = api.ProvingKey.keygen(params, cs, to_host(fixed), copies, tr)
    # pinned host copy of the witness for the end-to-end (host-buffer) path
    w.advice_pinned = w.advice_dev.cpu().pin_memory()
    w.advice_host = w.advice_pinned.numpy().view(np.uint64)
    w.h2d_bytes = w.advice_host.nbytes + sum(i.nbytes for i in w.instances)
    # compact host form (bit / byte / u64 columns) when every cell of the witness fits 64 bits
    w.compact = None
    if hasattr(w.circ, "advice_limbs"):
        try:
            w.compact = api.CompactAdvice([api.CompactAdvice.pack_canonical(c) for c in w.circ.advice_limbs])
        except ValueError:
            w.compact = None
    return w
